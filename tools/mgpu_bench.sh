#!/bin/bash
# usage: tools/mgpu_bench.sh N [extra bench flags]   -- one JSON line per run, kernel table on stderr
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N "$@" 2>/dev/null | grep '^{' | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    print('ms/step %.3f value %.4g host enqueue %.3f ms/step launches/step %.1f' % (d['ms_per_step'], d['value'], d.get('host_enqueue_ms_per_step', -1), d['gpu_launches'] / d['steps']))
    print(' '.join('%s=%.3f(x%.1f)' % (k, v['ms_per_step'], v['launches_per_step']) for k, v in d['kernels'].items()))
    print(d['config']['workload'])
"
