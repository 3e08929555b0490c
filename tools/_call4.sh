mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 4 --steps 20 --warmup 5 2>gpurun_out/bench_4gpu.err | grep '^{' > gpurun_out/r02_bench_c3_4gpu.json
grep -m3 "NCCL INFO" gpurun_out/bench_4gpu.err | cut -c1-200; grep -c "NCCL INFO" gpurun_out/bench_4gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c3_4gpu.json'))
print('C3 x4', d['value'], d['ms_per_step'], d['gpu_launches'], d['config']['workload'][-100:])
print(json.dumps(d.get('mgpu_parity'))[:400])
print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
print(d['e2e']['value'])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 4 --config C5 --steps 20 --warmup 5 2>gpurun_out/bench_c5.err | grep '^{' > gpurun_out/r02_bench_c5_4gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c5_4gpu.json'))
print('C5 x4', d['value'], d['ms_per_step'], d['gpu_launches'], d['config']['workload'][-100:])
print(json.dumps(d.get('mgpu_parity'))[:200])
print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
PY
