#!/bin/bash
# round 2, cycle p: k_solve with two polls in flight per lane (NB_SOLVE_POLL=2), parity under it and A/B on the same box
# (record of a measurement: the NB_SOLVE_POLL variants it selects were dropped from the source afterwards, see profiles/r02pq_solver_polling_experiments.txt)
mkdir -p gpurun_out
NB_SOLVE_POLL=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -3
for pv in 1 2 1 2; do
  NB_SOLVE_POLL=$pv python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-throughput-leg > gpurun_out/r02p_c2_poll$pv.json 2> gpurun_out/r02p_c2_poll$pv.err
  python -c "
import json; d=json.load(open('gpurun_out/r02p_c2_poll$pv.json')); print('poll $pv: c2', round(d['value'],1), round(d['ms_per_step'],4), 'solve launch ms', round(d['roofline']['avg_launch_ms'],4), 'e2e', round(d['e2e']['value'],1))"
done
NB_SOLVE_POLL=2 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline --no-throughput-leg > gpurun_out/r02p_c4_poll2.json 2> gpurun_out/r02p_c4_poll2.err
python -c "
import json; d=json.load(open('gpurun_out/r02p_c4_poll2.json')); print('poll 2: c4', round(d['value'],2), round(d['ms_per_step'],3), 'solve launch ms', round(d['roofline']['avg_launch_ms'],3))"
