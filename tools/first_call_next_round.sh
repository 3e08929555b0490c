#!/bin/bash
# First GPU call of the next round, one B200:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/first_call_next_round.sh'
# 1. the whole GPU suite WITHOUT -x (every failure of the files test_gpu_w0..w9 / x that have never run on
#    hardware is listed, not only the first), per-test durations;
# 2. the default bench line and the reference arm (is the default path where round 1 left it?);
# 3. step time of the new paths at C2 size (tools/bench_variants.py): PML, cross-deposition, antenna, external field.
# Outputs under gpurun_out/ (copy what is to be judged into profiles/).
mkdir -p gpurun_out
R=${ROUND:-r02}
timeout 1500 python -m pytest tests -m gpu -q -rfEs --durations=25 -p no:cacheprovider > gpurun_out/${R}_pytest_gpu_full.log 2>&1
tail -n 60 gpurun_out/${R}_pytest_gpu_full.log
python bench.py 2>gpurun_out/bench_default.err | grep '^{' > gpurun_out/${R}_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/${R}_bench_reference_arm.json
python tools/bench_variants.py > gpurun_out/${R}_bench_variants.json 2>gpurun_out/bench_variants.err
cut -c1-400 gpurun_out/${R}_bench_default.json; cat gpurun_out/${R}_bench_variants.json
