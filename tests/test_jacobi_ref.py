"""CPU checks of oracle/jacobi_ref.py (the restatement the throughput-mode solver is tested against) against the C oracle, which is
pinned bit-for-bit to the compiled reference: (1) with every body counted once the mass split must reproduce the reference's
effective-mass planes (nudge.cpp:4350-4561), (2) on a scene where no body has two contacts a Jacobi pass IS a Gauss-Seidel sweep,
so it must reproduce apply_impulses (nudge.cpp:4640-4855) up to the reference's rcpps/rsqrtps approximation error."""
import numpy as np
from nudge_b200 import scenes
from oracle import pyoracle, jacobi_ref as J


def _rows41(o):
    v = o.constraints_view()
    c = v["contact"]
    n = o.contacts.count
    first = np.full(n, -1, np.int64)
    for lane in range(len(c) - 1, -1, -1):      # first lane of every contact (leftover batches repeat lane 0)
        if c[lane] < n: first[c[lane]] = lane
    assert (first >= 0).all()
    a, b = v["a"][first].astype(np.int64), v["b"][first].astype(np.int64)
    P = np.zeros((41, n), np.float32)
    P[:39] = v["rows"][first].T
    P[39] = o.properties["mass_inverse"][a]; P[40] = o.properties["mass_inverse"][b]
    return P, v["states"][first].T.copy(), a, b


def _setup(o):
    o.collide(); o.apply_gravity_damping(); o.read_cached_impulses(); o.setup_contact_constraints()


def test_split_with_unit_counts_reproduces_the_reference_planes():
    o = pyoracle.OracleSim(scenes.demo_scene(120, 120, iterations=4, spread=3.0, height=12.0))
    for _ in range(60):
        o.step()
    _setup(o)
    P, st, a, b = _rows41(o)
    assert P.shape[1] > 200
    t = J.split_terms(P, a, b, np.ones(o.scene.n_bodies, np.int64))
    for k in ("NVTNI", "BIAS", "FC_X", "FC_Y", "FC_Z"):
        want = P[J.IX[k]].astype(np.float64)
        scale = np.abs(want).max()
        assert np.abs(t[k] - want).max() <= 2e-5 * scale, k


def test_jacobi_pass_equals_a_gauss_seidel_sweep_when_no_body_has_two_contacts():
    n = 36
    s = scenes.demo_scene(0, n, iterations=1)
    k = np.arange(n)
    s.transforms["position"][1:, 0] = (k % 6) * 4.0; s.transforms["position"][1:, 2] = (k // 6) * 4.0
    s.transforms["position"][1:, 1] = -10.0 + s.sphere_data["radius"] - 0.02      # resting on the ground box (top at y = -10), slightly sunk
    rng = np.random.default_rng(4)
    s.momentum["velocity"][1:] = rng.normal(size=(n, 3)) * 0.3; s.momentum["angular_velocity"][1:] = rng.normal(size=(n, 3)) * 0.3
    o = pyoracle.OracleSim(s)
    _setup(o)
    P, st, a, b = _rows41(o)
    assert P.shape[1] == n and len(np.unique(np.concatenate([a[a != 0], b[b != 0]]))) == n
    lin0, ang0 = o.momentum["velocity"].copy(), o.momentum["angular_velocity"].copy()
    o.apply_impulses()
    lin1, ang1, st1 = J.jacobi_pass(P, st, a, b, lin0, ang0)
    assert np.abs(lin1 - o.momentum["velocity"]).max() < 2e-3 and np.abs(ang1 - o.momentum["angular_velocity"]).max() < 2e-3
    assert np.abs(lin1 - lin0).max() > 1e-2       # the sweep did something
    _, st_ref, _, _ = _rows41(o)
    assert np.abs(st1 - st_ref).max() < 2e-3 * max(1.0, np.abs(st_ref).max())
