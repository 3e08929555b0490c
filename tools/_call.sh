python bench.py --config C2w --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2w', d['value'], d['ms_per_step'], d['ms_per_step_instrumented'], d['host_enqueue_ms_per_step']); print(d['e2e']['value'], d['e2e']['split_s']); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
python bench.py --config C2w --steps 56 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2w 56 steps', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'])"
timeout 300 python -m pytest tests/test_gpu_step.py tests/test_gpu_w4_scripts.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
