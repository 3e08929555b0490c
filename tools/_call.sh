timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_x_fullsize_properties.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for v in plain tiled; do
B2_GATHER_CUBIC=$v python bench.py --config C2c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2c $v', d['value'], d['ms_per_step']); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
done
