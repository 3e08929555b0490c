python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2', d['value'], d['ms_per_step']); print(d['e2e'])"
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_w4_scripts.py tests/test_gpu_w8_diags.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
