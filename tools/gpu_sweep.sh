#!/bin/bash
for t in 0 1 2 3 4; do
  DISO_BWD_TILE=$t python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-ref-cuda 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('tile', $t, 'step', round(d['ms_per_step'],3), 'mc_backward', d['kernels']['mc_backward']['ms'])"
done
