import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nudge_b200
from nudge_b200 import scenes
s = scenes.box_drop(70000)
g = nudge_b200.Sim(s)
rng = np.random.default_rng(1)
for n in [0, 1, 255, 256, 257, 1000, 151552, 151553, 435499, 1000000]:
    d = rng.integers(0, 5, n).astype(np.uint32)
    out, tot = g.device_scan(d)
    ref = np.concatenate([[0], np.cumsum(d)[:-1]]).astype(np.uint32) if n else d
    print("scan n=%d ok=%s total ok=%s" % (n, np.array_equal(out, ref), tot == int(d.sum())))
    for bits in (34, 40, 48):
        keys = rng.integers(0, 1 << bits, n, dtype=np.uint64)
        if n > 10: keys[: n // 3] &= np.uint64(0x1ffff)   # many small-hi keys like ground pairs
        vals = np.arange(n, dtype=np.uint32)
        k2, v2 = g.device_sort(keys, vals, 0, bits)
        order = np.argsort(keys, kind="stable")
        okk = np.array_equal(k2, keys[order]); okv = np.array_equal(v2, vals[order])
        print("  sort n=%d bits=%d keys ok=%s vals ok=%s" % (n, bits, okk, okv))
        if not okk:
            bad = np.nonzero(k2 != keys[order])[0]
            print("    bad rows", len(bad), bad[:10], bad[-5:])
            print("    got ", k2[bad[:6]]); print("    want", keys[order][bad[:6]])
