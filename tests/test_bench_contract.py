"""bench.py's contract with the driver, the part that runs without a GPU: `--impl reference` prints ONE JSON line on stdout with the keys
the driver reads (metric / unit / config of the repo's arm, the reference's own CPU implementation timed on host threads), and nothing else
on stdout.  Needs the prebuilt oracle/_ref/libnudge_ref_fast.so (skipped otherwise).  CPU only."""
import json, os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libnudge_ref_fast.so")):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "simulation steps/s" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] >= 3 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["config"] == "c2" and "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(d["ms_per_step"] * d["value"] / 1e3 - 1.0) < 1e-6


def test_bench_refuses_to_run_the_gpu_arm_without_a_gpu():
    """No CPU fallback: the repo's own arm must fail, loudly, when there is no CUDA device (it must never print a number)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and not [l for l in r.stdout.splitlines() if l.strip().startswith("{")], r.stdout[-500:]


def test_roofline_table_tool_reproduces_the_committed_table():
    """scripts/roofline_table.py on the committed ncu CSV and bench line regenerates profiles/r02f_roofline_c2.md (the per-kernel roofline
    table is an output of committed inputs, not hand-written)."""
    csv = os.path.join(ROOT, "profiles", "r02f_kernels_c2_metrics.csv"); line = os.path.join(ROOT, "profiles", "r02f_bench_c2.json")
    want = os.path.join(ROOT, "profiles", "r02f_roofline_c2.md")
    if not (os.path.exists(csv) and os.path.exists(line) and os.path.exists(want)):
        pytest.skip("profiles not present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "roofline_table.py"), csv, line, "--steps", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1000:]
    rows = lambda text: [l for l in text.splitlines() if l.startswith("| `")]
    got, ref = rows(r.stdout), rows(open(want).read())
    assert len(got) == len(ref) > 40
    assert got[0].startswith("| `k_solve`") and got == ref
