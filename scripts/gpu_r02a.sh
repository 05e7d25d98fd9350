# round 2, GPU cycle A: regression tests, all five BASELINE configs on one GPU, hand-off microbenchmark, per-kernel DRAM/L2 metrics
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/r02a_pytest.log; tail -15 gpurun_out/r02a_pytest.log
timeout 120 scripts/bin/handoff_latency > gpurun_out/r02a_handoff.txt 2>&1; cat gpurun_out/r02a_handoff.txt
for c in c2 c1 c3 c5 c4; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 $( [ $c != c2 ] && echo --no-cpu-baseline ) > gpurun_out/r02a_$c.json 2> gpurun_out/r02a_$c.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02a_$c.json'))
    print('$c', round(d['value'],1), 'steps/s', round(d['ms_per_step'],3), 'ms', d['config']['contacts'], 'contacts', {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4))
except Exception as e:
    print('$c failed', e); print(open('gpurun_out/r02a_$c.err').read()[-1500:])
PY
done
for c in c2 c4; do
  timeout 900 python bench.py --config $c --solver throughput --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_${c}_tp.json 2> gpurun_out/r02a_${c}_tp.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02a_${c}_tp.json'))
    print('$c throughput', round(d['value'],1), 'steps/s', round(d['ms_per_step'],3), 'ms', d['config']['contacts'], 'contacts', {k:round(v,3) for k,v in d['stage_ms'].items()}, 'frac', round(d['roofline']['frac'],4), 'launch ms', round(d['roofline']['avg_launch_ms'],4))
except Exception as e:
    print('$c throughput failed', e); print(open('gpurun_out/r02a_${c}_tp.err').read()[-1500:])
PY
done
NB_CUDA_PROFILER=staged timeout 900 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_jacobi -c 12 --csv --log-file gpurun_out/r02a_jacobi_c4.csv python bench.py --config c4 --solver throughput --presim 300 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo ncu_jacobi_rc=$?
NB_CUDA_PROFILER=1 timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r02a_kernels_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo ncu_rc=$?
ls -la gpurun_out | grep r02a
