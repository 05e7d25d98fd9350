# k_schedule with eight chunks in flight: parity suite, c2 / c4 / c5 parity lines, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6
for c in c2 c4 c5; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_$c.json 2> gpurun_out/r02i_$c.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02i_$c.json')); print('$c', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()})
except Exception as e: print('$c failed', e); print(open('gpurun_out/r02i_$c.err').read()[-800:])
PY
done
NB_CUDA_PROFILER=1 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02i_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
grep -E "k_schedule|k_solve" gpurun_out/r02i_launches_warm.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-80 | head -6
