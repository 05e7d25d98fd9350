"""CPU study for DESIGN.md §7(1): how many hand-offs of the solver's dependency chains would stay inside one thread block
if contacts were assigned to blocks by spatial cluster?  Uses the CPU oracle (test infrastructure, hence under tests/) on a settled
pile; no GPU.  Not a pytest module: run it by hand.

For every contact (in schedule order) the predecessor on each of its two bodies is an edge of the dependency graph.  A contact is
owned by the cluster of one of its bodies; an edge is "local" when both contacts have the same owner.  The script reports the
share of local edges and the length of the critical path when a local hand-off costs `t_local` and a remote one `t_remote`."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import numpy as np
from nudge_b200 import scenes
from oracle import pyoracle


def critical(a, b, owner, sweeps, t_local, t_remote, t_compute):
    last_t = {}; last_owner = {}
    local = remote = 0
    n = len(a)
    for w in range(sweeps):
        for i in range(n):
            t = 0.0
            for body in (a[i], b[i]):
                if not body: continue
                if body in last_t:
                    same = last_owner[body] == owner[i]
                    local += same; remote += (not same)
                    t = max(t, last_t[body] + (t_local if same else t_remote))
            t += t_compute
            for body in (a[i], b[i]):
                if body: last_t[body] = t; last_owner[body] = owner[i]
    return max(last_t.values()), local, remote


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8191
    settle = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    s = scenes.box_drop(n, iterations=8, seed=2)
    o = pyoracle.OracleSim(s)
    t0 = time.time()
    for _ in range(settle): o.step()
    o.collide(); o.apply_gravity_damping(); o.read_cached_impulses(); o.setup_contact_constraints()
    v = o.constraints_view()
    used = v["contact"] != 0xffffffff
    a = v["a"][used].astype(np.int64); b = v["b"][used].astype(np.int64)
    print("boxes %d, %d contacts in %d batches (settled %d steps, %.0f s)" % (n, len(a), v["batches"], settle, time.time() - t0))
    pos = o.transforms["position"] if "position" in o.transforms.dtype.names else None
    P = np.array([[t[0][0], t[0][1], t[0][2]] for t in o.transforms]) if pos is None else pos
    for cell in (3.0, 4.5, 6.0, 9.0):
        cid = np.floor(P / cell).astype(np.int64)
        key = (cid[:, 0] * 73856093) ^ (cid[:, 1] * 19349663) ^ (cid[:, 2] * 83492791)
        # owner of a contact: the cluster of its dynamic body with the smaller cluster key (body 0 = static world has none)
        ka = np.where(a > 0, key[a], np.iinfo(np.int64).max); kb = np.where(b > 0, key[b], np.iinfo(np.int64).max)
        owner = np.minimum(ka, kb)
        sizes = np.unique(owner, return_counts=True)[1]
        for tl in (0.15,):
            T, loc, rem = critical(a.tolist(), b.tolist(), owner.tolist(), 8, tl, 0.55, 0.30)
            T0, _, _ = critical(a.tolist(), b.tolist(), owner.tolist(), 8, 0.55, 0.55, 0.30)
            print("cell %.1f: %d clusters, contacts per cluster median %d max %d; local edges %.0f %%; critical path %.0f us vs %.0f us all-remote"
                  % (cell, len(sizes), int(np.median(sizes)), int(sizes.max()), 100.0 * loc / (loc + rem), T, T0))


if __name__ == "__main__":
    main()
