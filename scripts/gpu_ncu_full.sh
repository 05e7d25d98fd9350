# usage: bash scripts/gpu_ncu_full.sh <kernel-regex> <out-name> [launch-count]
NB_CUDA_PROFILER=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k "regex:$1" -c ${3:-2} -o gpurun_out/$2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/$2.log 2>&1
ls -la gpurun_out/$2.ncu-rep
