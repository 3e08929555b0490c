mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "pipe or golden" -p no:cacheprovider 2>&1 | tail -3
# ncu --set full of the hot kernels inside the bench command (steady state: skip the set-up launches)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_dht_tma|k_gather_push|k_deposit_mma" --launch-skip 60 -c 10 -o gpurun_out/r02_hot -f python bench.py --steps 4 --warmup 3 --preroll 8 --no-e2e --no-cpu-baseline > gpurun_out/ncu_hot.log 2>&1; tail -2 gpurun_out/ncu_hot.log
B2_GATHER_IMPL=pipe timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_gather_push" --launch-skip 10 -c 2 -o gpurun_out/r02_gather_pipe -f python bench.py --steps 4 --warmup 3 --preroll 8 --no-e2e --no-cpu-baseline > gpurun_out/ncu_gp.log 2>&1; tail -2 gpurun_out/ncu_gp.log
# launch list of the bench command (per-launch durations, cold-cache / serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --preroll 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
