"""Manual GPU bring-up script (run under gpurun): prints the host CPU, validates the LUT model, then runs stage parity."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nudge_b200
from nudge_b200 import scenes, abi
from oracle import pyoracle
from tests.parity_util import Report, compare_oracle_gpu_step, sync_oracle_from_gpu

def main():
    print(subprocess.run("lscpu | grep -E 'Model name|^CPU\\(s\\)|Flags' | cut -c1-400; nvidia-smi -L", shell=True, capture_output=True, text=True).stdout)
    which = sys.argv[1:] or ["small", "demo", "drop"]
    olib = pyoracle.load()
    first = True
    for w in which:
        if w == "small": s = scenes.demo_scene(100, 100, iterations=4, spread=2.0, height=20.0); steps = 25
        elif w == "demo": s = scenes.demo_scene(1024, 1024, iterations=8); steps = 12
        elif w == "drop": s = scenes.box_drop(3000, iterations=8); steps = 60
        elif w == "drop64k": s = scenes.box_drop(65536, iterations=8); steps = 2
        elif w == "drop8k": s = scenes.box_drop(8000, iterations=8); steps = 2
        o = pyoracle.OracleSim(s)
        g = nudge_b200.Sim(s, contact_capacity=o.cap, debug=True)
        if w in ("drop64k", "drop8k"):
            pre = 1500 if w == "drop64k" else 600
            t0 = time.time()
            for i in range(pre):
                g.step()
                if i % 100 == 99:
                    c = g.counts(); print("  presim step %d: contacts %d pairs %d levels %d overflow %d, %.2f ms/step" % (i, c.contacts, c.pairs, c.levels, c.overflow, 1e3 * (time.time() - t0) / (i + 1)))
            sync_oracle_from_gpu(o, g)
        if first:
            first = False
            print("lut model exact on this host:", g.lut_model_exact())
            rng = np.random.default_rng(0)
            x = rng.integers(0, 2**32, 1 << 17, dtype=np.uint64).astype(np.uint32).view(np.float32)
            for rs in (False, True):
                y = np.empty_like(x); (olib.nbo_rsqrt if rs else olib.nbo_rcp)(abi.ptr(x), abi.ptr(y), len(x))
                yd = g.device_rcp(x, rs)
                nan = np.isnan(y)
                same = (y.view(np.uint32) == yd.view(np.uint32)) | (nan & np.isnan(yd))
                print("device %s vs host instruction: %d / %d mismatches" % ("rsqrt" if rs else "rcp", int((~same).sum()), len(x)))
        t0 = time.time()
        for i in range(steps):
            rep = Report("%s step %d" % (w, i))
            ok = compare_oracle_gpu_step(o, g, rep)
            c = g.counts()
            print("%s step %d: contacts %d pairs %d batches %d levels %d overflow %d -> %s" % (w, i, c.contacts, c.pairs, c.batches, c.levels, c.overflow, "OK" if ok else "FAIL"))
            if not ok:
                print(rep)
                break
        print("%s: %.1fs, launches so far %d" % (w, time.time() - t0, g.launch_count()))
        g.close()

if __name__ == "__main__":
    main()
