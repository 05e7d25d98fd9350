"""The oracle against the committed golden vectors (generated from the unmodified reference by tests/golden/make_golden.py):
the reference's own known-answer tests (tests/main.cpp: contact counts 4 / 8 / 3 / 1 / 2, tag algebra) and a short
trajectory through all seven calls.  CPU only; needs neither /root/reference nor oracle/_ref."""
import os
import numpy as np
from nudge_b200 import scenes
from tests import golden_util as G


def test_known_answer_counts_in_fixture():
    g = G.load_box_cases()
    assert (g["count"] == g["expect"]).all()
    assert set(np.unique(g["expect"])) == {1, 2, 3, 4, 8}


def test_oracle_matches_reference_two_box_cases():
    from oracle import pyoracle
    g = G.load_box_cases()
    errs = []
    for k in range(len(g["count"])):
        o = pyoracle.OracleSim(G.scene_of(g, k), contact_capacity=64)
        o.collide()
        e = G.check_case(g, k, o.contacts_view())
        if e:
            errs.append(e)
    assert not errs, "\n".join(errs[:10])


def test_face_tags_are_unique_within_a_manifold():
    """tests/main.cpp:347-353: no duplicate feature tags inside one face-face manifold."""
    g = G.load_box_cases()
    for k in np.nonzero(g["family"] == 1)[0]:
        t = g["tags"][k, :8] & np.uint64(0xffffffff)
        assert len(np.unique(t)) == 8


def test_oracle_matches_reference_trajectory():
    from oracle import pyoracle
    t = np.load(os.path.join(G.HERE, "golden", "step_cases.npz"))
    s = scenes.demo_scene(48, 48, iterations=8, seed=int(t["seed"]), spread=2.0, height=12.0)
    o = pyoracle.OracleSim(s)
    for k in range(len(t["contacts"])):
        o.step()
        assert o.contacts.count == t["contacts"][k] and o.cache.count == t["cache"][k]
        assert np.array_equal(o.transforms.view(np.uint8), t["transforms"][k].view(np.uint8)), "transforms differ at step %d" % k
        assert np.array_equal(o.momentum.view(np.uint8), t["momentum"][k].view(np.uint8)), "momentum differs at step %d" % k
