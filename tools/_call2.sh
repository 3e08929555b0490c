mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench_2gpu.err | grep '^{' > gpurun_out/r02_bench_c3_2gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c3_2gpu.json'))
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['config']['workload'][-100:])
print(json.dumps(d.get('mgpu_parity'))[:300])
for k,v in d['kernels'].items(): print('   %-12s %.4f ms %.2f launches'%(k, v['ms_per_step'], v['launches_per_step']))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --config C5 --steps 10 --warmup 3 --no-parity 2>gpurun_out/bench_c5.err | grep '^{' > gpurun_out/r02_bench_c5_2gpu.json
tail -3 gpurun_out/bench_c5.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c5_2gpu.json'))
print('C5', d['value'], d['ms_per_step'], d['gpu_launches'], d['config']['workload'][-100:])
for k,v in d['kernels'].items(): print('   %-12s %.4f ms %.2f launches'%(k, v['ms_per_step'], v['launches_per_step']))
PY
