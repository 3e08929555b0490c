timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_w7_multi.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench_c3_2gpu.err | grep '^{' > gpurun_out/r02_bench_c3_2gpu.json
cut -c1-400 gpurun_out/r02_bench_c3_2gpu.json
