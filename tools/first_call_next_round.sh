#!/bin/bash
# First GPU call of a round, one B200:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/first_call_next_round.sh'
# 1. the whole GPU suite WITHOUT -x (every failure is listed, not only the first), per-test durations;
# 2. the default bench line and the reference arm as the driver runs them (--steps 20 --warmup 5);
# 3. the other BASELINE configs on one GPU (C1, C4 both particle loads).
# Outputs under gpurun_out/ (copy what is to be judged into profiles/).
mkdir -p gpurun_out
R=${ROUND:-r02}
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc; grep -m1 'model name' /proc/cpuinfo
timeout 1500 python -m pytest tests -m gpu -q -rfEs --durations=25 -p no:cacheprovider > gpurun_out/${R}_pytest_gpu_full.log 2>&1
tail -n 60 gpurun_out/${R}_pytest_gpu_full.log
python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/bench_ref.err | grep '^{' > gpurun_out/${R}_bench_reference_arm.json
cut -c1-600 gpurun_out/${R}_bench_reference_arm.json; tail -n 3 gpurun_out/bench_ref.err
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_default.err | grep '^{' > gpurun_out/${R}_bench_default.json
cut -c1-400 gpurun_out/${R}_bench_default.json; tail -n 3 gpurun_out/bench_default.err
for c in C1 C4; do
  python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_$c.err | grep '^{' > gpurun_out/${R}_bench_$c.json
  cut -c1-300 gpurun_out/${R}_bench_$c.json
done
