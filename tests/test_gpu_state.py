"""State files (nb_save_state / nb_load_state) and the headless replay tool tools/nb_replay.cpp: a dumped state,
loaded into a fresh context or replayed by the CLI in another process, continues bit-identically."""
import os, subprocess, tempfile
import numpy as np
import pytest
import nudge_b200
from nudge_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fnv1a(b):
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xffffffffffffffff
    return h


def test_state_roundtrip_and_headless_replay():
    import torch
    s = scenes.demo_scene(300, 300, iterations=8, spread=4.0, height=30.0)
    s.connections = np.zeros(3, scenes.PAIR32); s.connections["a"] = (1, 5, 9); s.connections["b"] = (2, 6, 10)
    side = torch.cuda.Stream()
    g = nudge_b200.Sim(s, stream=side.cuda_stream)
    for _ in range(60):
        g.step()
    d = tempfile.mkdtemp(prefix="nb_state_")
    p1, p2 = os.path.join(d, "a.bin"), os.path.join(d, "b.bin")
    g.save_state(p1)
    other = scenes.demo_scene(300, 300, iterations=8, seed=99); other.connections = np.zeros(3, scenes.PAIR32)   # different content, same capacities
    h = nudge_b200.Sim(other, stream=side.cuda_stream)
    h.load_state(p1)
    for _ in range(10):
        g.step(); h.step()
    g.download_bodies(); h.download_bodies(); g.download_cache(); h.download_cache()
    for name in ("transforms", "momentum", "idle"):
        assert getattr(g, name).tobytes() == getattr(h, name).tobytes(), name
    n = g.cache.count
    assert n == h.cache.count and g.cache_tags[:n].tobytes() == h.cache_tags[:n].tobytes() and g.cache_data[:n].tobytes() == h.cache_data[:n].tobytes()
    # the CLI, another process: same 10 steps from the same file
    exe = os.path.join(ROOT, "tools", "bin", "nb_replay")
    assert os.path.exists(exe), "tools/bin/nb_replay not built (nudge_b200/csrc/build.sh)"
    out = subprocess.run([exe, p1, "--steps", "10", "--iterations", "8", "--dump", p2], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    last = out.stdout.strip().splitlines()[-1].split()
    assert last[0] == "step" and last[1] == "9"
    assert int(last[last.index("xf") + 1], 16) == _fnv1a(g.transforms.tobytes())
    assert int(last[last.index("contacts") + 1]) == g.counts().contacts
    k = nudge_b200.Sim(other, stream=side.cuda_stream)
    k.load_state(p2)
    k.download_bodies()
    assert k.transforms.tobytes() == g.transforms.tobytes() and k.momentum.tobytes() == g.momentum.tobytes() and k.idle.tobytes() == g.idle.tobytes()
