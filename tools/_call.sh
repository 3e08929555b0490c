timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_w4_scripts.py tests/test_gpu_w6_acceptance.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for v in 1 0; do B2_OVERLAP_EB=$v python bench.py --config C2w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2w overlap=$v', d['value'], d['ms_per_step'], d['ms_per_step_instrumented']); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"; done
