#!/bin/bash
# round 2, cycle m: renderer read-back + demo trajectory tests, default bench line with the throughput-mode leg
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6
python bench.py > gpurun_out/r02m_bench_c2.json 2> gpurun_out/r02m_bench_c2.err; tail -c 300 gpurun_out/r02m_bench_c2.err
python -c "
import json
d=json.load(open('gpurun_out/r02m_bench_c2.json'))
print('c2', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))
t=d['throughput_mode']; print('throughput leg', round(t['value'],1), round(t['ms_per_step'],4), 'frac', round(t['roofline']['frac'],3), t['roofline']['avg_launch_ms'])
print({k:round(v,3) for k,v in d['stage_ms'].items()})"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NB_CUDA_PROFILER=1 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02m_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-throughput-leg > /dev/null 2>&1; echo rc=$?
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/r02m_launches_warm.csv') if l.startswith('"')]
rd=list(csv.DictReader(lines)); rd=rd[len(rd)//2:]
print([(re.sub(r"\(.*$","",r["Kernel Name"]).replace("void ","")[:14], round(float(r["Metric Value"])/1000,1)) for r in rd if "sort_coop" in r["Kernel Name"] or "k_solve" in r["Kernel Name"] or "k_schedule" in r["Kernel Name"]])
PY
