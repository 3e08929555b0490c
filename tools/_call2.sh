timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_w7_multi.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2 | tee gpurun_out/r02_pytest_gpu_2gpu.log
for v in 1 0; do
B2_OVERLAP_EB=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e 2>gpurun_out/bench_c3_2gpu.err | grep '^{' > gpurun_out/r02_bench_c3_2gpu_overlap$v.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c3_2gpu_overlap$v.json')); print('overlap=$v', d['value'], d['ms_per_step'], d['ms_per_step_instrumented'], d['mgpu_parity']['ok'])"
done
