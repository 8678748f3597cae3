#!/bin/bash
# solve time versus the number of sub-batches (development aid)
for wl in "$@"; do
  for ns in 1 2 4 8 16; do
    echo -n "$wl nsub=$ns: "
    DDPB_SUBBATCHES=$ns python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.3e ms/step %.2f e2e %.3e one-stream %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_timing']['ms_per_step_one_stream']), {k: round(v['ms_avg'], 3) for k, v in d['kernels'].items()})
"
  done
done
