/* TEST INFRASTRUCTURE ONLY — headless stand-in for <GL/glut.h> (see GL/gl.h next to it).  glutMainLoop does not open a window: it
 * hands control to demo_capture_loop() (oracle/demo_capture.cpp), which calls the demo's display and timer callbacks itself. */
#pragma once
#include "gl.h"
enum { GLUT_RGBA = 0, GLUT_DEPTH = 16, GLUT_DOUBLE = 2 };
struct GlutStubState { void (*display)(); void (*timer)(int); };
inline GlutStubState& glut_stub() { static GlutStubState s = {}; return s; }
inline void glutInit(int*, char**) {}
inline void glutInitDisplayMode(unsigned) {}
inline void glutInitWindowSize(int, int) {}
inline int glutCreateWindow(const char*) { return 1; }
inline void glutDisplayFunc(void (*f)()) { glut_stub().display = f; }
void demo_capture_before_simulate();   /* the demo's timer() re-arms itself right before it calls simulate() (example/main.cpp:330-334) */
inline void glutTimerFunc(unsigned, void (*f)(int), int) { glut_stub().timer = f; demo_capture_before_simulate(); }
inline void glutPostRedisplay() {}
inline void glutSwapBuffers() {}
inline void glutSolidCube(GLdouble) { gl_stub_emit(); }
inline void glutSolidSphere(GLdouble, GLint, GLint) { gl_stub_emit(); }
void demo_capture_loop();
inline void glutMainLoop() { demo_capture_loop(); }
