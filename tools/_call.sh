mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
export DHT_BENCH_REPS=20
for st in 0 1500 3000 6000; do echo "== stagger $st"; B2_DHT_STAGGER_NS=$st timeout 300 python tools/dht_bench.py --one 2>&1 | grep "whole\|EB" | grep -v C1 | cut -c1-130; done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/r02_d_bench_default.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_d_bench_default.json'))
print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])
print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
print(d['e2e'])
PY
for c in C2w C4t; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_$c.err | grep '^{' > gpurun_out/r02_bench_$c.json; tail -2 gpurun_out/bench_$c.err | cut -c1-300; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_$c.json'))
    print('$c', d['value'], d['ms_per_step'], d['config']['workload'][-60:]); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
except Exception as e: print('$c failed', e)
PY
done
