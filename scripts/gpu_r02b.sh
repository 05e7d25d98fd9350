# round 2, GPU cycle B: throughput-test details, full suite on the newest build, parity bench with the sort/overlap changes, remaining throughput configs,
# solver traffic captures, sanitizer
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_throughput.py -m gpu -q --timeout 300 2>&1 | tail -150 > gpurun_out/r02b_pytest_tp.log; tail -5 gpurun_out/r02b_pytest_tp.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_throughput.py 2>&1 | tail -60 > gpurun_out/r02b_pytest.log; tail -8 gpurun_out/r02b_pytest.log
run() { # name, args...
  name=$1; shift
  timeout 900 python bench.py "$@" --steps 20 --warmup 3 > gpurun_out/r02b_$name.json 2> gpurun_out/r02b_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02b_$name.json'))
    print('$name', round(d['value'],1), 'steps/s', round(d['ms_per_step'],3), 'ms', d['config']['contacts'], 'contacts', {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'launch ms', round(d['roofline']['avg_launch_ms'],4), d.get('dropin_seven_call_steps_per_s'))
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/r02b_$name.err').read()[-1500:])
PY
}
run c2
NB_OVERLAP=0 run c2_nooverlap --no-cpu-baseline
run c5_tp --config c5 --solver throughput --no-cpu-baseline
run c3_tp --config c3 --solver throughput --no-cpu-baseline
run c3 --config c3 --no-cpu-baseline
# solver traffic on the benchmarked states: dram bytes of k_jacobi_sweep (1M boxes) and k_solve (64k)
NB_CUDA_PROFILER=staged timeout 900 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_jacobi_sweep -c 6 --csv --log-file gpurun_out/r02b_jacobi_c4_traffic.csv python bench.py --config c4 --solver throughput --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_jacobi_c4_traffic.json 2>/dev/null; echo ncu_rc=$?
NB_CUDA_PROFILER=staged timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_jacobi_sweep -c 2 -o gpurun_out/r02b_jacobi_full -f python bench.py --config c4 --solver throughput --presim 400 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo ncu_full_rc=$?
NB_CUDA_PROFILER=1 timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r02b_kernels_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo ncu_rc=$?
bash scripts/sanitize.sh
ls -la gpurun_out | grep r02b
