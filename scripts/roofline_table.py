#!/usr/bin/env python
"""Per-kernel roofline table from an ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum, lts__t_bytes.sum
per launch) and the bench line of the same workload (for the element counts).

  python scripts/roofline_table.py gpurun_out/r02a_kernels_c2.csv gpurun_out/r02a_c2.json --steps 1 > profiles/r02_roofline.md

Algorithmic bytes per kernel are the figures of DESIGN.md section 2 / SURVEY.md section 8(d) (K colliders, P broadphase pairs,
C contacts, B bodies, A active bodies, I sweeps); kernels without a stated figure show their measured DRAM traffic only.
Times under ncu are serialised and (with --cache-control none) warm-cache: compare SHARES, not absolutes."""
import csv, json, re, sys, argparse
from collections import OrderedDict


def load_ncu(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    launches = OrderedDict()
    for r in rd:
        k = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)   # -> microseconds
        elif unit in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        k[name] = v
    return list(launches.values())


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    return name


def algorithmic(counts):
    K, P, C, B, A, I = (counts[k] for k in "KPCBAI")
    return [  # (regex, bytes per STEP, what)
        (r"^k_collider_world", 150.0 * K, "150 B/collider"),
        (r"^k_morton", 48.0 * K, "32 B AABB read + 12 B key/value + extent bins"),
        (r"^k_leaves", 60.0 * K, "sorted AABB rows"),
        (r"^k_grid_build", 44.0 * K, ""),
        (r"^k_grid_pairs", 32.0 * 27 * K / 8 + 8.0 * P, "candidate boxes 32 B each (L1/L2 hits mostly) + pair keys"),
        (r"^k_np_faces", 96.0 * P, "2 x (32 B transform + 16 B box) per pair"),
        (r"^k_np_clip", 96.0 * 0.6 * P + 52.0 * C, "surviving pairs + contacts parked"),
        (r"^k_np_emit", 2 * 52.0 * C, "contacts moved into place"),
        (r"^k_contact_compact", 2 * 52.0 * C, ""),
        (r"^k_build_rows", (52.0 + 180.0) * C + 96.0 * 2 * C, "contact 52 B read, 45 planes written, body gathers"),
        (r"^k_solve", (184.0 * C + 64.0 * A) * I, "184 B/contact + 64 B/active body per sweep, all sweeps in one launch"),
        (r"^k_jacobi_sweep", (184.0 * C + 64.0 * A) * (I + 1), "184 B/contact + 64 B/active body per pass (warm start + sweeps)"),
        (r"^k_jacobi_apply", 128.0 * A * (I + 1), "velocity + accumulator rows read and written"),
        (r"^k_update_impulses", (12.0 + 36.0 + 16.0) * C, ""),
        (r"^k_cache_merge", 24.0 * C * 2, ""),
        (r"^k_cache_lookup", 40.0 * C, ""),
        (r"^k_advance", 99.0 * A, "99 B/active body"),
        (r"^k_gravity_damping", 68.0 * A, ""),
        (r"^k_mw_in", 64.0 * B, ""), (r"^k_mw_out", 128.0 * B, ""),
        (r"^k_inertia", 80.0 * B, ""),
        (r"^k_sort_coop", None, "latency bound (grid barriers), see DESIGN.md 2.2"),
        (r"^k_schedule", None, "serial replay of the reference's scheduler, latency bound"),
    ]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ncu_csv"); ap.add_argument("bench_json")
    ap.add_argument("--steps", type=int, default=1, help="simulation steps covered by the capture")
    ap.add_argument("--peak", type=float, default=None)
    a = ap.parse_args()
    launches = load_ncu(a.ncu_csv)
    bench = json.load(open(a.bench_json))
    cfg = bench["config"]
    peak = a.peak or bench.get("roofline", {}).get("peak", 6570.6)
    counts = dict(K=cfg.get("colliders_per_gpu", cfg.get("bodies_per_gpu", 0)), P=cfg.get("broadphase_pairs", 0), C=cfg.get("contacts", 0),
                  B=cfg.get("bodies_per_gpu", 0), A=cfg.get("bodies_per_gpu", 1) - 1, I=cfg.get("solver_iterations", 8))
    alg = algorithmic(counts)
    per = OrderedDict()
    for l in launches:
        n = short(l["name"])
        d = per.setdefault(n, dict(n=0, us=0.0, dram=0.0, l2=0.0))
        d["n"] += 1; d["us"] += l.get("gpu__time_duration.sum", 0.0)
        d["dram"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0); d["l2"] += l.get("lts__t_bytes.sum", 0.0)
    total_us = sum(d["us"] for d in per.values())
    S = float(a.steps)
    print("# Per-kernel roofline, %s\n" % cfg.get("workload", ""))
    print("Source: `%s` (ncu: gpu__time_duration, dram__bytes_read/write, lts__t_bytes per launch; %d launches over %d step(s)); counts from `%s`: "
          "K=%d colliders, P=%d pairs, C=%d contacts, B=%d bodies, I=%d sweeps.  Peak = %.1f GB/s (measured copy bandwidth, MEASURED_PEAKS.json).  "
          "Times under ncu are serialised per launch: shares are meaningful, absolutes are not the step time.\n" %
          (a.ncu_csv, len(launches), a.steps, a.bench_json, counts["K"], counts["P"], counts["C"], counts["B"], counts["I"], peak))
    print("| kernel | launches/step | µs/step | share | DRAM MB/step | DRAM GB/s | L2 MB/step | algorithmic MB/step | algorithmic GB/s | frac of peak | note |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for n, d in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        us = d["us"] / S
        ab, note = None, ""
        for rx, b, what in alg:
            if re.search(rx, n):
                ab, note = b, what
                break
        dram_gbs = d["dram"] / S / (us * 1e-6) / 1e9 if us > 0 else 0.0
        if ab:
            gbs = ab / (us * 1e-6) / 1e9
            algs = "%.2f | %.1f | %.3f" % (ab / 1e6, gbs, gbs / peak)
        else:
            algs = "— | — | —"
        print("| `%s` | %.1f | %.1f | %.1f %% | %.2f | %.1f | %.2f | %s | %s |" % (n, d["n"] / S, us, 100.0 * d["us"] / total_us, d["dram"] / S / 1e6, dram_gbs, d["l2"] / S / 1e6, algs, note))
    print("\nTotal: %.1f µs of kernel time per step across %d launches/step." % (total_us / S, round(len(launches) / S)))


if __name__ == "__main__":
    main()
