#!/usr/bin/env python
"""Generates tests/golden/demo_frames.npz from the reference's own demo application, run headless in THIS container:

    python tests/golden/make_demo_golden.py            (needs /root/reference; builds oracle/_ref/demo_capture through oracle/Makefile)

oracle/demo_capture.cpp compiles /root/reference/example/main.cpp unmodified against recording GL stand-ins; the demo builds its scene from
rand() (1 ground box + 1024 boxes + 512 spheres, example/main.cpp:391-432) and every frame runs render() and simulate() (two sub-steps of
1/120 s with 20 iterations, example/main.cpp:274-334).  Recorded: the initial state (before the first simulate()), and at FRAMES (frame f = after f + 1 calls of simulate()) the body and collider arrays the demo owns, and the model
matrices its render() passed to glLoadMatrixf (example/main.cpp:224-268).  Used by tests/test_render_ref.py (oracle restatement of the
matrices), tests/test_gpu_render.py (nb_instance_matrices) and tests/test_gpu_parity.py (the demo's own trajectory, bit for bit)."""
import os, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FRAMES = (0, 40, 900)


def parse(path):
    buf = open(path, "rb").read()
    off, out = 0, {}
    def take(n, dt):
        nonlocal off
        a = np.frombuffer(buf, dt, n, off).copy()
        off += (a.nbytes + 3) & ~3
        return a
    while off < len(buf):
        magic, frame, nb, nbox, nsph, nmat = np.frombuffer(buf, "<u4", 6, off); off += 24
        assert magic == 0x4f4d444e
        k = "initial_" if frame == 0xffffffff else "f%d_" % frame
        out[k + "transforms"] = take(nb * 8, "<u4").reshape(nb, 8)       # raw words: position[3], body, rotation[4]
        out[k + "momentum"] = take(nb * 8, "<f4").reshape(nb, 8)
        out[k + "properties"] = take(nb * 4, "<f4").reshape(nb, 4)
        out[k + "idle"] = take(nb, "u1")
        out[k + "box_transforms"] = take(nbox * 8, "<u4").reshape(nbox, 8)
        out[k + "box_data"] = take(nbox * 4, "<f4").reshape(nbox, 4)
        out[k + "box_tags"] = take(nbox, "<u2")
        out[k + "sphere_transforms"] = take(nsph * 8, "<u4").reshape(nsph, 8)
        out[k + "sphere_data"] = take(nsph, "<f4")
        out[k + "sphere_tags"] = take(nsph, "<u2")
        out[k + "matrices"] = take(nmat * 16, "<f4").reshape(nmat, 16)
    return out


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    exe = os.path.join(ROOT, "oracle", "_ref", "demo_capture")
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "demo.bin")
        subprocess.check_call([exe], env=dict(os.environ, DEMO_CAPTURE_OUT=p, DEMO_CAPTURE_FRAMES=",".join(str(f) for f in FRAMES)), stdout=subprocess.DEVNULL)
        out = parse(p)
    out["frames"] = np.asarray(FRAMES, np.uint32)
    # the demo exactly as shipped (FTZ/DAZ on, example/main.cpp:338-339), compared at the last frame
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "demo_ftz.bin")
        subprocess.check_call([exe], env=dict(os.environ, DEMO_CAPTURE_OUT=p, DEMO_CAPTURE_FRAMES=str(FRAMES[-1]), DEMO_CAPTURE_KEEP_FTZ="1"), stdout=subprocess.DEVNULL)
        ftz = parse(p)
    # (identical on this trajectory - no denormal ever arises - so only that fact is stored: the recorded frames ARE the demo as shipped)
    out["ftz_daz_run_identical"] = np.asarray([np.array_equal(ftz["f%d_transforms" % FRAMES[-1]], out["f%d_transforms" % FRAMES[-1]])])
    # the collider arrays do not change between frames: keep one copy
    for f in FRAMES:
        for n in ("box_transforms", "box_data", "box_tags", "sphere_transforms", "sphere_data", "sphere_tags", "properties"):
            assert np.array_equal(out["f%d_%s" % (f, n)], out["initial_" + n])
            del out["f%d_%s" % (f, n)]
    del out["initial_matrices"]
    dst = os.path.join(HERE, "demo_frames.npz")
    np.savez_compressed(dst, **out)
    print(dst, os.path.getsize(dst), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
