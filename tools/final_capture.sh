#!/bin/bash
# Round-end measurement on one B200 (outputs under gpurun_out/, copy what is to be judged into profiles/):
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/final_capture.sh'
mkdir -p gpurun_out
R=${ROUND:-r02}
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc; grep -m1 'model name' /proc/cpuinfo
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -rfEs --durations=10 -p no:cacheprovider > gpurun_out/${R}_pytest_gpu_full.log 2>&1
tail -n 8 gpurun_out/${R}_pytest_gpu_full.log
python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/bench_ref.err | grep '^{' > gpurun_out/${R}_bench_reference_arm.json
cut -c1-300 gpurun_out/${R}_bench_reference_arm.json
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_default.err | grep '^{' > gpurun_out/${R}_bench_default.json
cut -c1-300 gpurun_out/${R}_bench_default.json
for c in C1 C4 C4t C2w C2c; do
  python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_$c.err | grep '^{' > gpurun_out/${R}_bench_$c.json
  cut -c1-200 gpurun_out/${R}_bench_$c.json
done
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --preroll 2 --no-e2e --no-cpu-baseline > /dev/null 2>&1
# ncu --set full of the hot kernels in steady state (DRAM traffic per launch -> tools/ncu_traffic.py)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_dht_tma|k_gather_push|k_deposit_mma|k_fft_pass|k_spectral" --launch-skip 110 -c 16 -o gpurun_out/${R}_hot -f python bench.py --steps 4 --warmup 3 --preroll 8 --no-e2e --no-cpu-baseline > gpurun_out/ncu_hot.log 2>&1; tail -1 gpurun_out/ncu_hot.log
