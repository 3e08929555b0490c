for L in 1 2 4 8; do
  echo "LANES=$L"
  B2_FFT_LANES=$L python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print(d['ms_per_step'], d['value'], {k: round(v, 3) for k, v in d.get('profile_ms_per_step', {}).items()} if 'profile_ms_per_step' in d else d.get('config'))
"
done
python -m pytest tests/test_gpu_step.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3
