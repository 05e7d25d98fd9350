#!/bin/bash
# round 2, cycle r: k_solve hands a body's 32-byte row over with ONE 256-bit access (NB_SOLVE_WIDE=1): whole GPU suite under it, A/B on the same box
mkdir -p gpurun_out
NB_SOLVE_WIDE=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
run() {
  env "${@:2}" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-throughput-leg > gpurun_out/r02r_c2_$1.json 2> gpurun_out/r02r_c2_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/r02r_c2_$1.json')); print('$1: c2', round(d['value'],1), round(d['ms_per_step'],4), 'solve launch ms', round(d['roofline']['avg_launch_ms'],4), 'e2e', round(d['e2e']['value'],1))"
}
run narrow NB_SOLVE_WIDE=0
run wide NB_SOLVE_WIDE=1
run narrow2 NB_SOLVE_WIDE=0
run wide2 NB_SOLVE_WIDE=1
