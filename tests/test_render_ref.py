"""The recorded run of the reference's own demo (tests/golden/demo_frames.npz) against the CPU restatements: the renderer's model
matrices (oracle/render_ref.py vs what example/main.cpp:224-268 handed to glLoadMatrixf) and the demo's trajectory (the widened oracle
stepping the demo's initial state vs the state the demo reached).  CPU only; needs neither /root/reference nor oracle/_ref."""
import numpy as np
from nudge_b200 import scenes as S
from oracle import render_ref
from tests import golden_util as G


def test_render_oracle_matches_the_demo_matrices():
    g = G.load_demo_frames()
    for f in g["frames"]:
        s = G.demo_scene(g, "f%d" % f)
        m = render_ref.instance_matrices(s.transforms, s.box_transforms, s.box_data["size"], s.sphere_transforms, s.sphere_data["radius"])
        want = g["f%d_matrices" % f]
        assert m.shape == want.shape == (s.n_colliders, 16)
        assert np.array_equal(m.view(np.uint32), want.view(np.uint32)), "frame %d: %d matrices differ" % (f, (m.view(np.uint32) != want.view(np.uint32)).any(axis=1).sum())


def test_render_oracle_is_a_rigid_transform_times_scale():
    """Independent of the fixture: M * (corner, 1) = position + R(q) (scale * corner) with R from float64 quaternion algebra."""
    rng = np.random.default_rng(5)
    s = S.demo_scene(16, 16, seed=9)
    s.transforms["rotation"][1:] = S._random_unit_quaternions(rng, s.n_bodies - 1)
    s.box_transforms["position"][:] = rng.normal(size=(s.n_boxes, 3)).astype(np.float32) * 0.1
    m = render_ref.instance_matrices(s.transforms, s.box_transforms, s.box_data["size"], s.sphere_transforms, s.sphere_data["radius"]).astype(np.float64).reshape(-1, 4, 4).transpose(0, 2, 1)
    def rot(q, v):
        u, w = q[:, :3], q[:, 3:4]
        return v + 2.0 * np.cross(u, np.cross(u, v) + w * v)
    corner = np.array([1.0, -1.0, 1.0])
    cx = np.concatenate([s.box_transforms, s.sphere_transforms]); b = s.transforms[cx["body"]]
    scale = np.concatenate([s.box_data["size"], np.repeat(s.sphere_data["radius"][:, None], 3, axis=1)]).astype(np.float64)
    bq = b["rotation"].astype(np.float64)
    local = cx["position"].astype(np.float64) + rot(cx["rotation"].astype(np.float64), scale * corner)
    want = b["position"].astype(np.float64) + rot(bq, local)
    got = (m @ np.append(corner, 1.0))[:, :3]
    assert np.abs(got - want).max() < 1e-4


def test_oracle_follows_the_demo_trajectory():
    """The widened restatement, started from the demo's initial state, reaches the demo's recorded frames bit for bit: frames 0, 40 and 900
    (1802 sub-steps of 20 iterations; by then the pile has settled and 566 of the 1536 bodies sleep, so islands, idle counters, the culled
    part of the contact cache and waking are all on the path)."""
    from oracle import pyoracle
    g = G.load_demo_frames()
    o = pyoracle.OracleSim(G.demo_scene(g, "initial"))
    done = 0
    for f in g["frames"]:
        for _ in range(G.demo_substeps(int(f)) - done):
            o.step()
        done = G.demo_substeps(int(f))
        assert np.array_equal(o.transforms.view(np.uint32).reshape(-1, 8), g["f%d_transforms" % f]), "transforms differ at frame %d" % f
        assert np.array_equal(o.momentum.view(np.uint32).reshape(-1, 8), g["f%d_momentum" % f].view(np.uint32)), "momentum differs at frame %d" % f
        assert np.array_equal(o.idle, g["f%d_idle" % f])


def test_recorded_frames_are_the_demo_as_shipped():
    """The frames were recorded with denormals kept (the parity convention); the demo as shipped runs with FTZ/DAZ on (example/main.cpp:338-339).
    make_demo_golden.py ran both and stored whether the last frame differs: it does not, so the fixture is the unmodified demo's own output."""
    g = G.load_demo_frames()
    assert bool(g["ftz_daz_run_identical"][0])
