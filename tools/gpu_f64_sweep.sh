#!/bin/bash
for c in 50 58 65 72 86; do DISO_CARVEOUT_BWD=$c python bench.py --dtype f64 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('f64 carveout', $c, 'step', round(d['ms_per_step'],3), 'mc_backward', d['kernels']['mc_backward']['ms'])"; done
