"""Pins the widened CPU restatement (oracle/nudge_oracle.cpp) against the UNMODIFIED reference compiled in place
(oracle/_ref), bit for bit, at every stage of the step.  CPU only."""
import numpy as np
import pytest
from nudge_b200 import scenes
from tests.conftest import needs_ref
from tests.parity_util import Report, compare_ref_oracle_step


def _run(scene, steps):
    from oracle import pyref, pyoracle
    r = pyref.RefSim(scene); o = pyoracle.OracleSim(scene, contact_capacity=r.cap)
    for i in range(steps):
        rep = Report("%s step %d" % (scene.name, i))
        assert compare_ref_oracle_step(r, o, rep), str(rep)
    return r, o


@needs_ref
def test_small_mixed_scene_every_stage():
    _run(scenes.demo_scene(100, 100, iterations=4, spread=2.0, height=20.0), 30)


@needs_ref
def test_demo_scene_config0():
    """BASELINE config 0: 1024 boxes + 1024 spheres + ground, 8 iterations."""
    _run(scenes.demo_scene(1024, 1024, iterations=8), 4)


@needs_ref
def test_rotated_box_drop():
    _run(scenes.box_drop(1500, iterations=8), 12)


@needs_ref
def test_sleeping_islands_and_culled_cache():
    """Forces part of the scene asleep (idle counter 0xff) so that both island passes, sleeping pairs and the culled
    cache entries (nudge.cpp:3674-3700, 3973-4000, 4064-4101) are exercised."""
    from oracle import pyref, pyoracle
    s = scenes.demo_scene(120, 120, iterations=4, spread=6.0, height=6.0, seed=11)
    r = pyref.RefSim(s); o = pyoracle.OracleSim(s, contact_capacity=r.cap)
    for i in range(40):
        rep = Report("settle %d" % i)
        assert compare_ref_oracle_step(r, o, rep), str(rep)
    rng = np.random.default_rng(5)
    sleepy = rng.random(s.n_bodies) < 0.8
    r.idle[sleepy] = 0xff; o.idle[sleepy] = 0xff
    r.momentum["velocity"][sleepy] = 0; o.momentum["velocity"][sleepy] = 0
    r.momentum["angular_velocity"][sleepy] = 0; o.momentum["angular_velocity"][sleepy] = 0
    seen_sleeping = 0
    for i in range(6):
        rep = Report("sleep %d" % i)
        assert compare_ref_oracle_step(r, o, rep), str(rep)
        seen_sleeping = max(seen_sleeping, r.contacts.sleeping_count)
    assert seen_sleeping > 0, "scenario never produced sleeping pairs"


@needs_ref
def test_staged_calls_equal_fused_reference_step():
    from oracle import pyref
    s = scenes.demo_scene(64, 64, iterations=8, spread=2.0, height=10.0)
    a = pyref.RefSim(s); b = pyref.RefSim(s)
    for _ in range(20):
        a.step(); b.step_staged()
    assert np.array_equal(a.transforms.view(np.uint8), b.transforms.view(np.uint8))
    assert np.array_equal(a.momentum.view(np.uint8), b.momentum.view(np.uint8))


@needs_ref
def test_random_scenes_differential_fuzz_seeded():
    """A bounded, seeded run of tests/fuzz_oracle_vs_ref.py (random scenes, iterations, connections, sleep): every stage bit for bit."""
    from tests import fuzz_oracle_vs_ref as F
    rng = np.random.default_rng(77)
    for k in range(120):
        s = F.random_scene(rng)
        err = F.run_scene(s, rng, int(rng.integers(5, 40)))
        assert err is None, "scene %d (%s): %s" % (k, s.name, err[:600])
