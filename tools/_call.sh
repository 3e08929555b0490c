mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "transforms or dht" -p no:cacheprovider 2>&1 | tail -3
export DHT_BENCH_REPS=20
B2_DHT_IMPL=tma timeout 300 python tools/dht_bench.py --one 2>&1 | grep -v "C1" | cut -c1-150 | tee gpurun_out/r02_dht_bench_e.jsonl
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_plasma_wave.py tests/test_gpu_x_config_shapes.py tests/test_gpu_w4_scripts.py tests/test_gpu_w1_pml_cross.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_default.err | grep '^{' > gpurun_out/r02_b_bench_default.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_b_bench_default.json'))
print(d['value'], d['ms_per_step'], d['roofline'], d['e2e']['value'])
for k,v in d['kernels'].items(): print('   ',k, {a:round(b,4) for a,b in v.items()})
PY
