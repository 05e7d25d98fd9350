# round-end evidence on one GPU: both bench arms, ncu launch lists (warm and cold caches), one full capture of the solver
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 300 gpurun_out/final_ref.err
NB_CUDA_PROFILER=1 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
NB_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_cold.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
NB_CUDA_PROFILER=staged ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_solve -c 4 -o gpurun_out/final_solve -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_solve.log 2>&1
ls -la gpurun_out/ | grep final
