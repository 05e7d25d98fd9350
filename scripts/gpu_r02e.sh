# round 2, GPU cycle E: throughput kernel, one vs two shared stages, on the 1M-box and the 256k-brick scenes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_throughput.py -m gpu -q --timeout 300 2>&1 | tail -5
run() { # name, env, args...
  name=$1; envs=$2; shift 2
  env $envs timeout 900 python bench.py "$@" --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_$name.json 2> gpurun_out/r02e_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02e_$name.json'))
    print('$name', round(d['value'],1), 'steps/s', round(d['ms_per_step'],3), 'ms', d['config']['contacts'], 'contacts', 'solve', round(d['stage_ms']['apply_impulses'],3), 'frac', round(d['roofline']['frac'],4), 'launch ms', round(d['roofline']['avg_launch_ms'],4))
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/r02e_$name.err').read()[-1500:])
PY
}
run c4_s1 NB_JACOBI_STAGES=1 --config c4 --solver throughput
run c4_s2 NB_JACOBI_STAGES=2 --config c4 --solver throughput
run c5_s1 NB_JACOBI_STAGES=1 --config c5 --solver throughput
run c5_s2 NB_JACOBI_STAGES=2 --config c5 --solver throughput
run c3_s2 NB_JACOBI_STAGES=2 --config c3 --solver throughput
run c2_s2 NB_JACOBI_STAGES=2 --solver throughput
