# round 2, final single-GPU evidence: full GPU test suite, both bench arms on the default config, launch lists (warm and cold caches),
# per-kernel DRAM/L2 metrics, full captures of the two solver kernels, throughput-mode line on the 1M-box scene with its DRAM traffic
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/r02f_pytest.log; tail -4 gpurun_out/r02f_pytest.log
python bench.py > gpurun_out/r02f_bench_c2.json 2> gpurun_out/r02f_bench_c2.err; tail -c 300 gpurun_out/r02f_bench_c2.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err
python bench.py --config c4 --solver throughput --no-cpu-baseline > gpurun_out/r02f_bench_c4_throughput.json 2> gpurun_out/r02f_bench_c4_throughput.err
python bench.py --config c4 --no-cpu-baseline --steps 10 > gpurun_out/r02f_bench_c4_parity.json 2> gpurun_out/r02f_bench_c4_parity.err
python - <<'PY'
import json
for f in ("c2","reference","c4_throughput","c4_parity"):
    try:
        d=json.load(open("gpurun_out/r02f_bench_%s.json"%f)); print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
NB_CUDA_PROFILER=1 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
NB_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_cold.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
NB_CUDA_PROFILER=1 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r02f_kernels_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
NB_CUDA_PROFILER=staged ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_solve -c 2 -o gpurun_out/r02f_solve_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
NB_CUDA_PROFILER=staged ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_jacobi_sweep -s 2 -c 2 -o gpurun_out/r02f_jacobi_full -f python bench.py --config c4 --solver throughput --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out | grep r02f
