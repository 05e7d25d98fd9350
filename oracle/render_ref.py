"""CPU restatement of the demo's per-collider model matrix (TEST INFRASTRUCTURE ONLY; tests/ imports it).

Follows /root/reference/example/main.cpp: rotation = quaternion_concat(body, collider) (:53-58), position = quaternion_transform(body.rotation,
collider.position) + body.position (:60-72, :238-242), matrix(scale, rotation, position) (:74-110), scale = box half extents (:232) or the
sphere radius three times (:254), boxes first, then spheres (:224-268).  float32 throughout, one IEEE operation per source operation (the
pinned build is -ffp-contract=off), so the result is bit-comparable.  Pinned against the matrices the unmodified demo hands to glLoadMatrixf:
tests/golden/demo_frames.npz (oracle/demo_capture.cpp), tests/test_render_ref.py."""
import numpy as np


def instance_matrices(body_xf, box_xf, box_size, sphere_xf, sphere_radius):
    """body_xf / box_xf / sphere_xf: TRANSFORM arrays (nudge_b200.scenes); box_size [n, 3]; sphere_radius [m].  Returns [n + m, 16] float32."""
    f = np.float32
    cx = np.concatenate([box_xf, sphere_xf])
    s = np.concatenate([np.asarray(box_size, f).reshape(-1, 3), np.repeat(np.asarray(sphere_radius, f).reshape(-1, 1), 3, axis=1)])
    b = body_xf[cx["body"]]
    a0, a1, a2, a3 = (b["rotation"][:, k] for k in range(4))
    b0, b1, b2, b3 = (cx["rotation"][:, k] for k in range(4))
    q0 = b0*a3 + a0*b3 + a1*b2 - a2*b1
    q1 = b1*a3 + a1*b3 + a2*b0 - a0*b2
    q2 = b2*a3 + a2*b3 + a0*b1 - a1*b0
    q3 = a3*b3 - a0*b0 - a1*b1 - a2*b2
    c0, c1, c2 = (cx["position"][:, k] for k in range(3))
    t0 = a1*c2 - a2*c1; t1 = a2*c0 - a0*c2; t2 = a0*c1 - a1*c0
    t0 = t0 + t0; t1 = t1 + t1; t2 = t2 + t2
    p0 = c0 + a3*t0 + a1*t2 - a2*t1
    p1 = c1 + a3*t1 + a2*t0 - a0*t2
    p2 = c2 + a3*t2 + a0*t1 - a1*t0
    p0 = p0 + b["position"][:, 0]; p1 = p1 + b["position"][:, 1]; p2 = p2 + b["position"][:, 2]
    kx = q0 + q0; ky = q1 + q1; kz = q2 + q2
    xx = kx*q0; yy = ky*q1; zz = kz*q2; xy = kx*q1; xz = kx*q2; yz = ky*q2; sx = kx*q3; sy = ky*q3; sz = kz*q3
    one = f(1.0)
    m = np.zeros((len(cx), 16), f)
    m[:, 0] = (one - yy - zz) * s[:, 0]; m[:, 1] = (xy + sz) * s[:, 0]; m[:, 2] = (xz - sy) * s[:, 0]
    m[:, 4] = (xy - sz) * s[:, 1]; m[:, 5] = (one - xx - zz) * s[:, 1]; m[:, 6] = (yz + sx) * s[:, 1]
    m[:, 8] = (xz + sy) * s[:, 2]; m[:, 9] = (yz - sx) * s[:, 2]; m[:, 10] = (one - xx - yy) * s[:, 2]
    m[:, 12] = p0; m[:, 13] = p1; m[:, 14] = p2; m[:, 15] = one
    return m
