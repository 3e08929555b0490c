mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "transforms or dht" -p no:cacheprovider 2>&1 | tail -2
DHT_BENCH_REPS=20 timeout 300 python tools/dht_bench.py --one 2>&1 | grep -v C1 | cut -c1-130
timeout 600 ncu --set full --clock-control none -k regex:"k_dht_tma" --launch-skip 24 -c 2 -o gpurun_out/r02_dht2 -f python bench.py --steps 4 --warmup 3 --preroll 8 --no-e2e --no-cpu-baseline > gpurun_out/ncu_dht2.log 2>&1; tail -1 gpurun_out/ncu_dht2.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
