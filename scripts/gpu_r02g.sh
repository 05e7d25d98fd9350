# source-level profile of the cooperative sort (which phase costs what) and of k_schedule
mkdir -p gpurun_out
NB_CUDA_PROFILER=staged ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_sort_coop -c 5 -o gpurun_out/r02g_sort_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
NB_CUDA_PROFILER=staged ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_schedule -c 1 -o gpurun_out/r02g_sched_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out | grep r02g
