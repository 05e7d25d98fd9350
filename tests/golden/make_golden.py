"""Generates tests/golden/box_box_cases.npz by running the UNMODIFIED reference (oracle/_ref) on the two-box
configurations of the reference's own known-answer tests (/root/reference/tests/main.cpp):

  family 0  box_box_test_case_0            tests/main.cpp:130-200   -> 4 contacts
  family 1  box_box_face_face_tags_0       tests/main.cpp:202-357   -> 8 contacts
  family 2  box_box_face_face_tags_1       tests/main.cpp:359-514   -> 3 contacts
  family 3  box_box_edge_edge_tags         tests/main.cpp:516-667   -> 1 contact
  family 4  box_box_faces_share_tags       tests/main.cpp:669-834   -> 2 contacts

Each family sweeps the 4096 orientation combinations; every 29th combination is kept (plus the reversed collider
order).  Run from the repo root in the build container:  python tests/golden/make_golden.py
The fixture stores inputs and the reference's outputs, so tests need neither /root/reference nor oracle/_ref."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from nudge_b200 import scenes as S
from oracle import pyref

f32 = np.float32
PI = f32(3.14159265)


def rotate_axis_angle(r, ax, ay, az, angle):
    """tests/main.cpp:101-128 in float32."""
    angle = f32(angle)
    s = f32(np.sin(f32(angle * f32(0.5)), dtype=f32)); c = f32(np.cos(f32(angle * f32(0.5)), dtype=f32))
    a = np.array([f32(ax) * s, f32(ay) * s, f32(az) * s, c], f32)
    f = f32(1.0) / np.sqrt(f32(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]), dtype=f32)
    a = (a * f).astype(f32)
    b = r.copy()
    r[0] = b[0] * a[3] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1]
    r[1] = b[1] * a[3] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2]
    r[2] = b[2] * a[3] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0]
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]


def quat(seq):
    r = np.array([0, 0, 0, 1], f32)
    for (ax, ay, az, ang) in seq:
        rotate_axis_angle(r, ax, ay, az, ang)
    return r


def base(ax, ay, az):
    return [(0, 0, 1, PI * f32(0.5) * az), (0, 1, 0, PI * f32(0.5) * ay), (1, 0, 0, PI * f32(0.5) * ax)]


def cases():
    out = []
    # family 0: literal transform of tests/main.cpp:144-145, ground box 200x10x200 on body 0 at y=-10
    out.append(dict(family=0, expect=4, size=[(200, 10, 200), (0.5, 1.0, 0.25)], cpos=[(0, -10, 0), (0, 0, 0)], crot=[(0, 0, 0, 1), (0, 0, 0, 1)], cbody=[0, 1],
                    bpos=(-10.3300056, -0.0125209205, 20.0851059), brot=(0.0419846289, -0.296176672, -0.954125523, -0.0126985274)))
    for i in range(0, 4096, 29):
        ax, ay, az, bx, by, bz = [(i >> s) & 3 for s in (0, 2, 4, 6, 8, 10)]
        s2 = [(1.125, 1.125, 1.125)] * 2
        A, B = base(ax, ay, az), base(bx, by, bz)
        y45 = (0, 1, 0, PI * f32(0.25)); z45 = (0, 0, 1, PI * f32(0.25))
        # family 1 (tests/main.cpp:245-270): note the reversed order applies the tilt to the *other* box's chain
        out.append(dict(family=1, expect=8, size=s2, cpos=[(0, 1, 0), (0, -1, 0)], crot=[quat(A + [y45]), quat(B + [(1, 0, 0, PI * f32(1e-4))])], cbody=[0, 1]))
        out.append(dict(family=1, expect=8, size=s2, cpos=[(0, -1, 0), (0, 1, 0)], crot=[quat(B), quat(A + [(1, 0, 0, PI * f32(1e-4)), y45])], cbody=[1, 0]))
        # family 2 (tests/main.cpp:398-436)
        out.append(dict(family=2, expect=3, size=s2, cpos=[(0, 1, 2), (0, -1, 0)], crot=[quat(A + [y45]), quat(B + [(1, 0, 0, -PI * f32(1e-5))])], cbody=[0, 1]))
        out.append(dict(family=2, expect=3, size=s2, cpos=[(0, -1, 0), (0, 1, 2)], crot=[quat(B + [(1, 0, 0, PI * f32(1e-5))]), quat(A + [y45])], cbody=[1, 0]))
        # family 3 (tests/main.cpp:556-596)
        out.append(dict(family=3, expect=1, size=s2, cpos=[(2.2, 2.2, 0), (0, 0, 0)], crot=[quat(A + [(1, 0, 0, -PI * f32(1e-4)), y45, z45]), quat(B)], cbody=[0, 1]))
        out.append(dict(family=3, expect=1, size=s2, cpos=[(0, 0, 0), (2.2, 2.2, 0)], crot=[quat(B + [(1, 0, 0, PI * f32(1e-4))]), quat(A + [y45, z45])], cbody=[1, 0]))
        # family 4 (tests/main.cpp:709-757): +tilt and -tilt pick different faces of box 0
        out.append(dict(family=4, expect=2, size=s2, cpos=[(2.0, 0.1, 0), (0, 0, 0)], crot=[quat(A + [y45, (0, 1, 0, PI * f32(1e-3))]), quat(B)], cbody=[0, 1]))
        out.append(dict(family=4, expect=2, size=s2, cpos=[(2.0, 0.1, 0), (0, 0, 0)], crot=[quat(A + [y45, (0, 1, 0, -PI * f32(1e-3))]), quat(B)], cbody=[0, 1]))
    return out


def scene_of(c):
    s = S.Scene(2, 2, 0)
    s.box_tags[:] = (0, 1)
    s.box_data["size"][:] = np.asarray(c["size"], f32)
    s.box_transforms["position"][:] = np.asarray(c["cpos"], f32)
    s.box_transforms["rotation"][:] = np.asarray(c["crot"], f32)
    s.box_transforms["body"][:] = c["cbody"]
    if "bpos" in c:
        s.transforms["position"][1] = c["bpos"]; s.transforms["rotation"][1] = c["brot"]
    return s


def main():
    cs = cases()
    n = len(cs)
    arr = dict(family=np.zeros(n, np.int32), expect=np.zeros(n, np.int32), size=np.zeros((n, 2, 3), f32), cpos=np.zeros((n, 2, 3), f32), crot=np.zeros((n, 2, 4), f32),
               cbody=np.zeros((n, 2), np.uint32), bpos=np.zeros((n, 3), f32), brot=np.zeros((n, 4), f32), count=np.zeros(n, np.int32),
               contacts=np.zeros((n, 8), S.CONTACT), bodies=np.zeros((n, 8, 2), np.uint32), tags=np.zeros((n, 8), np.uint64))
    bad = 0
    for k, c in enumerate(cs):
        s = scene_of(c)
        r = pyref.RefSim(s, contact_capacity=128, arena_mb=4)
        r.collide()
        v = r.contacts_view()
        m = v["count"]
        arr["family"][k] = c["family"]; arr["expect"][k] = c["expect"]; arr["size"][k] = s.box_data["size"]; arr["cpos"][k] = s.box_transforms["position"]
        arr["crot"][k] = s.box_transforms["rotation"]; arr["cbody"][k] = s.box_transforms["body"]; arr["bpos"][k] = s.transforms["position"][1]; arr["brot"][k] = s.transforms["rotation"][1]
        arr["count"][k] = m; arr["contacts"][k, :m] = v["data"]; arr["bodies"][k, :m, 0] = v["bodies"]["a"]; arr["bodies"][k, :m, 1] = v["bodies"]["b"]; arr["tags"][k, :m] = v["tags"]
        if m != c["expect"]:
            bad += 1
    print("cases", n, "count != the reference test's expectation:", bad)
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "box_box_cases.npz"), **arr)

    # A short trajectory of a mixed scene through all seven calls (the reference's tests never leave collide()).
    s = S.demo_scene(48, 48, iterations=8, seed=7, spread=2.0, height=12.0)
    r = pyref.RefSim(s)
    steps = 24
    xf = np.zeros((steps, s.n_bodies), S.TRANSFORM); mom = np.zeros((steps, s.n_bodies), S.MOMENTUM); cnt = np.zeros(steps, np.int32)
    cache_n = np.zeros(steps, np.int32); tagsum = np.zeros(steps, np.uint64)
    for k in range(steps):
        r.step_staged()
        xf[k] = r.transforms; mom[k] = r.momentum; cnt[k] = r.contacts.count; cache_n[k] = r.cache.count
        tagsum[k] = np.bitwise_xor.reduce(r.cache_tags[:r.cache.count]) if r.cache.count else 0
    np.savez_compressed(os.path.join(here, "step_cases.npz"), seed=7, transforms=xf, momentum=mom, contacts=cnt, cache=cache_n, tag_xor=tagsum)
    print("trajectory contacts per step:", cnt)


if __name__ == "__main__":
    main()
