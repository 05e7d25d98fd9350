python -m pytest tests -m gpu -q -x 2>&1 | tail -3
NB_SORT=legacy python -m pytest tests/test_gpu_prims.py tests/test_gpu_parity.py -m gpu -q -x -k "prims or sort or 64k or golden" 2>&1 | tail -2
for mode in coop legacy; do
NB_SORT=$mode python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_$mode.json 2>gpurun_out/b_$mode.err
python -c "
import json,sys;d=json.load(open('gpurun_out/b_$mode.json'));print('$mode',d['value'],d['ms_per_step'],d['stage_ms'] if 'stage_ms' in d else '',d['gpu_launches'],d['e2e']['value'])"
done
NB_CUDA_PROFILER=1 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo done
