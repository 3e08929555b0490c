#!/bin/bash
# Round-end evidence on one B200: GPU test suite, default bench line, reference arm, ncu launch list,
# ncu --set full capture of the particle kernels and the Hankel GEMM.  Outputs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r01_pytest_gpu.log
python bench.py 2>gpurun_out/bench_default.err | grep '^{' > gpurun_out/r01_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/r01_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/r01_ncu_launches.csv \
    python bench.py --steps 8 --warmup 3 --preroll 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_deposit_mma|k_gather_push|k_dht' -s 42 -c 8 -f -o gpurun_out/r01_final \
    python bench.py --steps 8 --warmup 3 --preroll 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/r01_pytest_gpu.log; cat gpurun_out/r01_bench_default.json | cut -c1-600; ls -la gpurun_out
