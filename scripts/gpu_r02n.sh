#!/bin/bash
# round 2, cycle n: scheduler inputs bucket-major (coalesced loads for the replay warp)
# (record of a measurement: the bucket-major layout it measured was reverted)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-throughput-leg > gpurun_out/r02n_c2.json 2> gpurun_out/r02n_c2.err
python -c "
import json; d=json.load(open('gpurun_out/r02n_c2.json')); print('c2', round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['stage_ms'].items()})"
NB_CUDA_PROFILER=1 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02n_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-throughput-leg > /dev/null 2>&1; echo rc=$?
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/r02n_launches_warm.csv') if l.startswith('"')]
rd=list(csv.DictReader(lines)); rd=rd[len(rd)//2:]
print([(re.sub(r"\(.*$","",r["Kernel Name"]).replace("void ","")[:14], round(float(r["Metric Value"])/1000,1)) for r in rd if "sort_coop" in r["Kernel Name"] or "k_solve" in r["Kernel Name"] or "k_sched" in r["Kernel Name"]])
PY
