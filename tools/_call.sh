timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_plasma_wave.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for impl in pipe tiled; do
B2_GATHER_IMPL=$impl python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$impl', d['value'], d['ms_per_step']); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
done
python bench.py --config C4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4', d['value'], d['ms_per_step']); print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
