# round 2, GPU cycle D: throughput kernel with one shared stage + registers (3 CTAs/SM), study of its impulses against parity mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_throughput.py -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/r02d_pytest_tp.log; tail -5 gpurun_out/r02d_pytest_tp.log
run() { # name, args...
  name=$1; shift
  timeout 900 python bench.py "$@" --steps 20 --warmup 3 > gpurun_out/r02d_$name.json 2> gpurun_out/r02d_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02d_$name.json'))
    print('$name', round(d['value'],1), 'steps/s', round(d['ms_per_step'],3), 'ms', d['config']['contacts'], 'contacts', {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'launch ms', round(d['roofline']['avg_launch_ms'],4))
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/r02d_$name.err').read()[-1500:])
PY
}
run c4_tp --config c4 --solver throughput --no-cpu-baseline
run c2_tp --solver throughput --no-cpu-baseline
run c5_tp --config c5 --solver throughput --no-cpu-baseline
run c3_tp --config c3 --solver throughput --no-cpu-baseline
NB_CUDA_PROFILER=staged timeout 900 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_jacobi_sweep -c 4 --csv --log-file gpurun_out/r02d_jacobi_c4_traffic.csv python bench.py --config c4 --solver throughput --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_jacobi_c4_traffic.json 2>/dev/null; echo ncu_rc=$?
NB_CUDA_PROFILER=staged timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_jacobi_sweep -c 2 -o gpurun_out/r02d_jacobi_full -f python bench.py --config c4 --solver throughput --presim 400 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo ncu_full_rc=$?
timeout 900 python tests/study_throughput_vs_parity.py > gpurun_out/r02d_throughput_vs_parity.txt 2>&1; cat gpurun_out/r02d_throughput_vs_parity.txt
ls -la gpurun_out | grep r02d
