"""User constraint rows on the device (nb_upload_constraint_rows, SURVEY.md section 8 f1: the hook of example/main.cpp:316) against the
sequential host loop an application would write there (oracle/rows_ref.py), and a hanging chain of ball joints as a physical check."""
import numpy as np
import pytest
import nudge_b200
from nudge_b200 import scenes
from oracle import rows_ref

pytestmark = pytest.mark.gpu


def _quat_rotate(q, v):
    x, y, z, w = q
    t = 2.0 * np.cross(q[:3], v)
    return v + w * t + np.cross(q[:3], t)


def ball_joint_rows(xf, a, b, anchor_a, anchor_b, dt, beta=0.2):
    """Three rows holding the world-space anchors of bodies a and b together (b = 0: a point fixed in the world)."""
    pa, qa = xf["position"][a].astype(np.float64), xf["rotation"][a].astype(np.float64)
    ra = _quat_rotate(qa, np.asarray(anchor_a, np.float64))
    if b:
        pb, qb = xf["position"][b].astype(np.float64), xf["rotation"][b].astype(np.float64)
        rb = _quat_rotate(qb, np.asarray(anchor_b, np.float64)); wb = pb + rb
    else:
        rb = np.zeros(3); wb = np.asarray(anchor_b, np.float64)
    err = (pa + ra) - wb
    rows = np.zeros(3, nudge_b200.ROW)
    for k in range(3):
        e = np.zeros(3); e[k] = 1.0
        rows[k]["a"], rows[k]["b"] = a, b
        rows[k]["lin_a"] = e; rows[k]["ang_a"] = np.cross(ra, e)
        rows[k]["lin_b"] = -e; rows[k]["ang_b"] = -np.cross(rb, e)
        rows[k]["bias"] = beta / dt * err[k]
        rows[k]["lo"], rows[k]["hi"] = -np.inf, np.inf
    return rows


def test_rows_equal_the_sequential_host_loop():
    """No contacts (bodies far apart, far above the ground): the solver stage is the user rows alone, applied after each of the 5 sweeps."""
    n = 40
    s = scenes.demo_scene(n, 0, iterations=5, spread=200.0, height=10.0, seed=3)
    s.transforms["position"][1:, 1] += 500.0
    rng = np.random.default_rng(1)
    s.transforms["rotation"][1:] = scenes._random_unit_quaternions(rng, n)
    s.momentum["velocity"][1:] = rng.normal(size=(n, 3)); s.momentum["angular_velocity"][1:] = rng.normal(size=(n, 3))
    s.gravity = np.float32(0.0); s.damping = np.float32(0.0)
    g = nudge_b200.Sim(s)
    m = 150
    rows = np.zeros(m, nudge_b200.ROW)
    rows["a"] = rng.integers(0, n + 1, m); rows["b"] = rng.integers(0, n + 1, m)
    same = rows["a"] == rows["b"]; rows["b"][same] = (rows["a"][same] % n) + 1
    for k in ("lin_a", "ang_a", "lin_b", "ang_b"):
        rows[k] = rng.normal(size=(m, 3))
    rows["bias"] = rng.normal(size=m) * 0.3
    rows["lo"] = np.where(rng.random(m) < 0.3, 0.0, -np.inf); rows["hi"] = np.where(rng.random(m) < 0.2, 0.5, np.inf)
    rows["impulse"] = np.where(rows["lo"] == 0.0, 0.1, rng.normal(size=m) * 0.1); rows["softness"] = np.where(rng.random(m) < 0.5, 0.01, 0.0)
    g.upload_constraint_rows(rows)
    want = rows.copy()
    lin, ang = s.momentum["velocity"].astype(np.float64), s.momentum["angular_velocity"].astype(np.float64)
    I = rows_ref.world_inverse_inertia(s.transforms["rotation"], s.properties["inertia_inverse"])
    minv = s.properties["mass_inverse"].astype(np.float64)
    rows_ref.apply_rows(want, lin, ang, minv, I, warm=True)
    for _ in range(5):
        rows_ref.apply_rows(want, lin, ang, minv, I)
    g.collide(); g.apply_gravity_damping(); g.read_cached_impulses(); g.setup_contact_constraints(); g.apply_impulses(5)
    assert g.counts().contacts == 0
    g.download_bodies()
    got = g.download_constraint_rows(m)
    scale = max(1.0, np.abs(lin).max(), np.abs(ang).max())
    assert np.abs(g.momentum["velocity"] - lin).max() < 2e-5 * scale and np.abs(g.momentum["angular_velocity"] - ang).max() < 2e-5 * scale
    assert np.abs(got["impulse"] - want["impulse"]).max() < 2e-5 * max(1.0, np.abs(want["impulse"]).max())
    # nb_step (graph replay on a created stream) takes the same path
    import torch
    side = torch.cuda.Stream()
    h = nudge_b200.Sim(s, stream=side.cuda_stream)
    h.upload_constraint_rows(rows)
    h.step(); g.update_cached_impulses(); g.write_cached_impulses(); g.advance()
    h.download_bodies(); g.download_bodies()
    assert h.momentum.tobytes() == g.momentum.tobytes() and h.transforms.tobytes() == g.transforms.tobytes()
    g.upload_constraint_rows(np.zeros(0, nudge_b200.ROW))        # removing the rows restores the plain step
    g.step_staged()


def test_hanging_chain_of_ball_joints_stays_connected_among_contacts():
    """A 10-link chain hung from the world by ball joints swings into a pile of loose boxes: the anchors stay together (joint error
    bounded) while ordinary contacts are solved around it; the joint bodies are connected for the island pass (example/main.cpp:285)."""
    links = 10
    s = scenes.demo_scene(links + 60, 0, iterations=12, spread=3.0, height=6.0, seed=5)
    s.box_data["size"][1:links + 1] = (0.2, 0.5, 0.2)
    s.properties[1:links + 1] = scenes._box_props(s.box_data["size"][1:links + 1])
    top = np.array([6.0, 8.0, 0.0])
    for k in range(links):                                    # laid out horizontally: it will swing down
        s.transforms["position"][1 + k] = top + np.array([0.5 + 1.0 * k, 0.0, 0.0])
        s.transforms["rotation"][1 + k] = (0.0, 0.0, np.sin(np.pi / 4), np.cos(np.pi / 4))   # long axis along x
    s.connections = np.zeros(links - 1, scenes.PAIR32); s.connections["a"] = 1 + np.arange(links - 1); s.connections["b"] = 2 + np.arange(links - 1)
    g = nudge_b200.Sim(s)
    dt = float(s.time_step)
    worst = 0.0
    for step in range(240):
        g.download_bodies()
        rows = [ball_joint_rows(g.transforms, 1, 0, (0.0, 0.5, 0.0), top, dt)]
        for k in range(1, links):
            rows.append(ball_joint_rows(g.transforms, 1 + k, k, (0.0, 0.5, 0.0), (0.0, -0.5, 0.0), dt))
        g.upload_constraint_rows(np.concatenate(rows))
        g.step_staged()
        if step > 20:
            g.download_bodies()
            for k in range(1, links):
                wa = g.transforms["position"][1 + k] + _quat_rotate(g.transforms["rotation"][1 + k].astype(np.float64), np.array([0.0, 0.5, 0.0]))
                wb = g.transforms["position"][k] + _quat_rotate(g.transforms["rotation"][k].astype(np.float64), np.array([0.0, -0.5, 0.0]))
                worst = max(worst, float(np.linalg.norm(wa - wb)))
    assert worst < 0.08, worst
    assert np.isfinite(g.transforms["position"]).all() and g.counts().overflow == 0
    end = g.transforms["position"][links]
    assert end[1] < top[1] - 0.5 * links                     # the chain hangs below its support now
