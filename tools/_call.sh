mkdir -p gpurun_out
R=r02
timeout 1500 python -m pytest tests -m gpu -q -rfEs --durations=10 -p no:cacheprovider > gpurun_out/${R}_pytest_gpu_full.log 2>&1
tail -n 25 gpurun_out/${R}_pytest_gpu_full.log
for impl in legacy pipe; do
B2_GATHER_IMPL=$impl python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$impl', d['value'], d['ms_per_step']); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
done
