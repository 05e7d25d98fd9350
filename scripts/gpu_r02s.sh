#!/bin/bash
# round 2, cycle s: the 256-bit hand-off as the default: smoke + one bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_c2.json 2> gpurun_out/r02s_c2.err
python -c "
import json; d=json.load(open('gpurun_out/r02s_c2.json')); print('c2', round(d['value'],1), round(d['ms_per_step'],4), 'solve', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'tp', round(d['throughput_mode']['value'],1))"
