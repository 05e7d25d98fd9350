#!/bin/bash
# round 2, cycle q: k_solve with a per-lane poll throttle (NB_SOLVE_POLL=3): parity under it, A/B on the same box for three lane-hop settings
# (record of a measurement: the NB_SOLVE_POLL variants it selects were dropped from the source afterwards, see profiles/r02pq_solver_polling_experiments.txt)
mkdir -p gpurun_out
NB_SOLVE_POLL=3 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -3
run() { # label, env...
  env "${@:2}" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-throughput-leg > gpurun_out/r02q_c2_$1.json 2> gpurun_out/r02q_c2_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/r02q_c2_$1.json')); print('$1: c2', round(d['value'],1), round(d['ms_per_step'],4), 'solve launch ms', round(d['roofline']['avg_launch_ms'],4))"
}
run base NB_SOLVE_POLL=1
run lane150 NB_SOLVE_POLL=3 NB_SOLVE_LANE_HOP_NS=150
run lane300 NB_SOLVE_POLL=3 NB_SOLVE_LANE_HOP_NS=300
run lane600 NB_SOLVE_POLL=3 NB_SOLVE_LANE_HOP_NS=600
run base2 NB_SOLVE_POLL=1
