timeout 300 python -m pytest tests/test_gpu_step.py tests/test_gpu_w4_scripts.py tests/test_gpu_w6_acceptance.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
python bench.py --config C2w --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_C2w.err | grep "^{" > gpurun_out/r02_bench_C2w.json; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_C2w.json')); print('C2w', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'])"
