"""Study (not a test): how far are the throughput-mode impulses from the parity-mode (= reference) impulses?

From ONE settled state of the 65,536-box pile, one step is solved with N sweeps in each mode; the contact impulses (what
update_cached_impulses writes, nudge.cpp:4857-4884) are compared contact by contact, and against parity mode with 256 sweeps as the
"converged" solution of the same contact problem.  Then each mode runs 200 further steps and the piles are compared.
  python tests/study_throughput_vs_parity.py [boxes] > profiles/r02_throughput_vs_parity.txt      (needs a GPU)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nudge_b200
from nudge_b200 import scenes


def snapshot(g):
    g.download_bodies(); g.download_cache()
    n = g.cache.count
    return dict(t=g.transforms.copy(), m=g.momentum.copy(), i=g.idle.copy(), n=n, tags=g.cache_tags[:n].copy(), f=g.cache_features[:n].copy(), d=g.cache_data[:n].copy())


def restore(g, s):
    g.transforms[:] = s["t"]; g.momentum[:] = s["m"]; g.idle[:] = s["i"]
    n = s["n"]; g.cache_tags[:n] = s["tags"]; g.cache_features[:n] = s["f"]; g.cache_data[:n] = s["d"]; g.cache.count = n
    g.upload_bodies(); g.upload_cache()


def solve_once(g, s, mode, sweeps):
    restore(g, s)
    g.set_solver_mode(mode)
    g.collide(); g.apply_gravity_damping(); g.read_cached_impulses(); g.setup_contact_constraints(); g.apply_impulses(sweeps); g.update_cached_impulses()
    g.download_contacts()
    n = g.contacts.count
    imp = g.debug("impulses", scenes.IMPULSE)["impulse"][:n].astype(np.float64)
    nrm = g.contact_data["normal"][:n].astype(np.float64)
    return imp, (imp * nrm).sum(1), g.contact_tags[:n].copy(), g.contact_features[:n].copy()


def main():
    boxes = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    scene = scenes.box_drop(boxes, iterations=8)
    g = nudge_b200.Sim(scene)
    for _ in range(900):
        g.step_staged()
    s = snapshot(g)
    ref, ref_n, tags, feats = solve_once(g, s, "parity", 256)
    weight = float((9.82 / scene.properties["mass_inverse"][1:]).sum() * float(scene.time_step))
    print("scene: %d boxes settled 900 steps, %d contacts; weight x dt of the pile = %.1f" % (boxes, len(ref), weight))
    print("impulses of ONE step from identical state; reference = parity mode (the reference's Gauss-Seidel order) with 256 sweeps")
    print("%-11s %7s %14s %14s %16s %16s" % ("mode", "sweeps", "sum normal J", "vs converged", "rel L2 vs conv.", "rel L2 vs parity@same"))
    par = {}
    for mode in ("parity", "throughput"):
        for sweeps in (8, 16, 32, 64, 128):
            imp, jn, t2, f2 = solve_once(g, s, mode, sweeps)
            assert np.array_equal(t2, tags) and np.array_equal(f2, feats)
            if mode == "parity":
                par[sweeps] = imp
            d_conv = np.linalg.norm(imp - ref) / np.linalg.norm(ref)
            d_par = np.linalg.norm(imp - par[sweeps]) / np.linalg.norm(par[sweeps])
            print("%-11s %7d %14.1f %13.1f%% %16.4f %16.4f" % (mode, sweeps, jn.sum(), 100.0 * jn.sum() / ref_n.sum(), d_conv, d_par))
    print()
    print("200 further steps in each mode (8 sweeps per step), from the same settled state:")
    for mode in ("parity", "throughput"):
        restore(g, s); g.set_solver_mode(mode)
        for _ in range(200):
            g.step_staged()
        g.collide(); g.download_contacts(); g.download_bodies()
        n = g.contacts.count
        pen = g.contact_data["penetration"][:n]
        v = g.momentum["velocity"][1:]
        print("%-11s contacts %7d  mean penetration %.4f  max %.3f  mean height %.4f  mean |v| %.4f" % (mode, n, pen.mean(), pen.max(), g.transforms["position"][1:, 1].mean(), np.linalg.norm(v, axis=1).mean()))


if __name__ == "__main__":
    main()
