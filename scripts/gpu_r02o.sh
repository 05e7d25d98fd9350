#!/bin/bash
# round 2, cycle o: nb_upload_bodies sends momentum + properties on a copy stream (under collide); e2e before/after on the same box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
for ov in 1 0; do
  NB_COPY_OVERLAP=$ov python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-throughput-leg > gpurun_out/r02o_c2_overlap$ov.json 2> gpurun_out/r02o_c2_overlap$ov.err
  python -c "
import json; d=json.load(open('gpurun_out/r02o_c2_overlap$ov.json')); print('overlap $ov: c2', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), round(1e3/d['e2e']['value'],4), 'ms')"
done
