timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fft_z" -p no:cacheprovider 2>&1 | tail -2
python tools/fft_sizes.py 2>&1 | grep Nz
timeout 300 python -m pytest tests/test_gpu_x_config_shapes.py -m gpu -q -x -k "c2_shape" -p no:cacheprovider 2>&1 | tail -2
for sp in 4 8 16; do python bench.py --steps 32 --warmup 5 --sort-period $sp --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('sort_period $sp', d['value'], d['ms_per_step']); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"; done
