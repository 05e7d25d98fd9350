/* TEST INFRASTRUCTURE ONLY — a recording stand-in for <GL/gl.h> so that the reference's demo (example/main.cpp, unmodified, compiled
 * where it lies) builds and runs headless: oracle/demo_capture.cpp.  Only what the demo calls; everything is a no-op except
 * glMatrixMode / glLoadMatrixf, which remember the model matrices the demo's render() hands to GL. */
#pragma once
#include <stddef.h>
#include <string.h>
typedef int GLint; typedef unsigned GLenum; typedef unsigned GLbitfield; typedef float GLfloat; typedef double GLdouble;
enum { GL_DEPTH_TEST = 1, GL_NORMALIZE, GL_LIGHTING, GL_LIGHT0, GL_VIEWPORT, GL_PROJECTION, GL_MODELVIEW, GL_AMBIENT, GL_DIFFUSE, GL_POSITION,
       GL_COLOR_BUFFER_BIT = 0x4000, GL_DEPTH_BUFFER_BIT = 0x100 };
struct GlStubState { GLenum mode; float current[16]; float* frame; size_t count, capacity; };
inline GlStubState& gl_stub() { static GlStubState s = {}; return s; }
inline void gl_stub_emit() {  /* a draw call: the current model matrix belongs to one collider */
	GlStubState& s = gl_stub();
	if (s.frame && s.count < s.capacity) memcpy(s.frame + 16 * s.count, s.current, sizeof(s.current));
	++s.count;
}
inline void glEnable(GLenum) {}
inline void glClearColor(GLfloat, GLfloat, GLfloat, GLfloat) {}
inline void glClear(GLbitfield) {}
inline void glGetIntegerv(GLenum, GLint* v) { v[0] = 0; v[1] = 0; v[2] = 1024; v[3] = 600; }
inline void glMatrixMode(GLenum m) { gl_stub().mode = m; }
inline void glLoadMatrixf(const GLfloat* m) { if (gl_stub().mode == GL_MODELVIEW) memcpy(gl_stub().current, m, sizeof(float) * 16); }
inline void glLoadIdentity() {}
inline void glTranslatef(GLfloat, GLfloat, GLfloat) {}
inline void glLightfv(GLenum, GLenum, const GLfloat*) {}
