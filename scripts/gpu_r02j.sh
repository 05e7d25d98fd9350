mkdir -p gpurun_out
NB_CUDA_PROFILER=staged ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_schedule -c 1 -o gpurun_out/r02j_sched_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
