# one GPU development cycle: parity tests, then the bench under a few settings, then a warm-cache launch list
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_$name.json 2>gpurun_out/b_$name.err
  python -c "
import json,sys;d=json.load(open('gpurun_out/b_$name.json'));print('$name',round(d['value'],1),round(d['ms_per_step'],4),{k:round(v,3) for k,v in d['stage_ms'].items()},d['gpu_launches'],round(d['e2e']['value'],1))"
}
run default NB_X=0
for spec in "$@"; do run "$(echo $spec | tr '= ' '__')" $spec; done
NB_CUDA_PROFILER=1 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo done
