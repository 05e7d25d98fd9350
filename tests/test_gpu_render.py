"""Renderer read-back (nb_instance_matrices) and the reference's own demo as a golden trajectory: the fixture
tests/golden/demo_frames.npz was recorded from /root/reference/example/main.cpp running unmodified and headless (oracle/demo_capture.cpp).
Bit-exact: the kernel performs the demo's float32 operations in the demo's order."""
import numpy as np
import pytest
import nudge_b200
from nudge_b200 import scenes
from oracle import render_ref
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def test_instance_matrices_match_what_the_demo_hands_to_gl():
    g = G.load_demo_frames()
    for f in g["frames"]:
        sim = nudge_b200.Sim(G.demo_scene(g, "f%d" % f))
        m = sim.instance_matrices()
        want = g["f%d_matrices" % f]
        assert m.shape == want.shape
        assert np.array_equal(m.view(np.uint32), want.view(np.uint32)), "frame %d: %d matrices differ" % (f, (m.view(np.uint32) != want.view(np.uint32)).any(axis=1).sum())


def test_instance_matrices_device_destination_and_collider_offsets():
    import torch
    rng = np.random.default_rng(11)
    s = scenes.demo_scene(700, 333, iterations=8, spread=5.0, height=40.0)
    s.box_transforms["position"][1:] = rng.normal(size=(s.n_boxes - 1, 3)).astype(np.float32) * 0.2      # colliders off their bodies' centres
    s.box_transforms["rotation"][1:] = scenes._random_unit_quaternions(rng, s.n_boxes - 1)
    s.sphere_transforms["position"][:] = rng.normal(size=(s.n_spheres, 3)).astype(np.float32) * 0.2
    side = torch.cuda.Stream()
    sim = nudge_b200.Sim(s, stream=side.cuda_stream)
    for _ in range(25):
        sim.step()
    sim.download_bodies()
    want = render_ref.instance_matrices(sim.transforms, s.box_transforms, s.box_data["size"], s.sphere_transforms, s.sphere_data["radius"])
    host = sim.instance_matrices()
    assert np.array_equal(host.view(np.uint32), want.view(np.uint32))
    with torch.cuda.stream(side):
        dev = torch.full((s.n_colliders + 3, 16), -1.0, dtype=torch.float32, device="cuda")
        n = sim.instance_matrices(device_ptr=dev.data_ptr(), capacity=len(dev))
    side.synchronize()
    assert n == s.n_colliders
    got = dev.cpu().numpy()
    assert np.array_equal(got[:n].view(np.uint32), want.view(np.uint32)) and (got[n:] == -1.0).all()
    with pytest.raises(RuntimeError):
        sim.instance_matrices(device_ptr=dev.data_ptr(), capacity=n - 1)


def test_gpu_follows_the_reference_demo_trajectory():
    """The demo's initial state stepped on the GPU (nb_step, 1/120 s, 20 iterations, gravity 9.82, damping 0.25 as example/main.cpp:274-305)
    reaches the frames the unmodified demo recorded, bit for bit."""
    g = G.load_demo_frames()
    sim = nudge_b200.Sim(G.demo_scene(g, "initial"))
    done = 0
    for f in (0, 40):       # frame 900 of the fixture (1802 sub-steps, sleeping islands) is the CPU oracle's pin: tests/test_render_ref.py
        for _ in range(G.demo_substeps(int(f)) - done):
            sim.step()
        done = G.demo_substeps(int(f))
        sim.download_bodies()
        assert sim.counts().overflow == 0
        assert np.array_equal(sim.transforms.view(np.uint32).reshape(-1, 8), g["f%d_transforms" % f]), "transforms differ at frame %d" % f
        assert np.array_equal(sim.momentum.view(np.uint32).reshape(-1, 8), g["f%d_momentum" % f].view(np.uint32)), "momentum differs at frame %d" % f
        assert np.array_equal(sim.idle, g["f%d_idle" % f])
        m = sim.instance_matrices()
        assert np.array_equal(m.view(np.uint32), g["f%d_matrices" % f].view(np.uint32))
