#!/bin/bash
# sweep environment settings over the 512^3 bench: gpu_sweep.sh "VAR=a" "VAR=b OTHER=c" ...
for setting in "$@"; do
  env $setting python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); k=d['kernels']
print('$setting', 'step', round(d['ms_per_step'],3), ' '.join('%s=%.3f'%(n.replace('dmc_','d').replace('mc_','m').replace('emit_',''), x['ms']) for n,x in k.items()))"
done
