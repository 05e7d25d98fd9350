# compute-sanitizer passes over a small scene: memcheck (global/shared out-of-bounds, misaligned, leaks of the
# relaxed 128-bit token words and bulk copies) and racecheck (shared-memory hazards: sort tiles, scheduler slots, TMA stage ring).
# usage: bash scripts/sanitize.sh   (on the GPU box; writes gpurun_out/sanitize_*.log)
mkdir -p gpurun_out
cat > /tmp/nb_sanitize_case.py <<'PY'
import numpy as np, nudge_b200
from nudge_b200 import scenes
s = scenes.demo_scene(96, 96, iterations=4, spread=2.0, height=8.0)
g = nudge_b200.Sim(s)
for _ in range(6): g.step_staged()
g.set_solver_mode("throughput")
for _ in range(4): g.step_staged()
c = g.counts(); print("contacts", c.contacts, "overflow", c.overflow)
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 env PYTHONPATH=$PWD python /tmp/nb_sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
