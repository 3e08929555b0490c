#!/bin/bash
# one short C2 bench line (kernel table) + the kernel / step parity tests
python bench.py --steps ${1:-60} --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print(d['ms_per_step'], d['value']); print(' '.join('%s=%.3f' % (k, v['ms_per_step']) for k, v in d['kernels'].items()))
"
