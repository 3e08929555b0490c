python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"k_dht_tma|k_gather_push|k_deposit_mma|k_fft_pass|k_spectral" --launch-skip 110 -c 16 -o gpurun_out/r02_hot -f python bench.py --steps 4 --warmup 3 --preroll 8 --no-e2e --no-cpu-baseline > gpurun_out/ncu_hot.log 2>&1; tail -1 gpurun_out/ncu_hot.log
