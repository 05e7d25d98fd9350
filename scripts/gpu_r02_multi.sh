# round 2, multi-GPU cycle: usage  bash scripts/gpu_r02_multi.sh <N> <tag> [bench args...]   (under gpurun --gpus N)
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
run() { # name, args...
  name=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_$name.json'))
    c=d['config']
    print('$name N=$N', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'scene steps/s', round(c['scene_steps_per_s'],1), 'other', c['other_transport'], round(c['other_transport_scene_steps_per_s'],1),
          'ghosts/gpu', c['ghost_bodies_per_gpu'], 'export', c['exchanged_rows_per_gpu_per_sweep'], 'graph' if 'graph' in c['step_call'] else 'plain', 'parity', d.get('parity_check'), 'e2e', round(d['e2e']['value'],1), 'reshard ms', round(c['reshard_ms_host_side_untimed'],1), 'ovf', c['overflow_flags'])
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/${TAG}_$name.err').read()[-3000:])
PY
}
for spec in "$@"; do
  name=$(echo "$spec" | cut -d: -f1); args=$(echo "$spec" | cut -d: -f2-)
  run $name $args --steps 20 --warmup 3
done
ls -la gpurun_out | grep ${TAG}
