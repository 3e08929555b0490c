timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2', d['value'], d['ms_per_step'], d['ms_per_step_instrumented']); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
python bench.py --config C4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4', d['value'], d['ms_per_step']); print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
