for c in C2w C4t; do python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$c', d['value'], d['ms_per_step']); print(d['e2e']['value'], d['e2e']['split_s'])"; done
