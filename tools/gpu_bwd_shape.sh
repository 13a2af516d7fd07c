#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
bash tools/gpu_sweep.sh DISO_CARVEOUT_BWD=50 DISO_CARVEOUT_BWD=58 DISO_CARVEOUT_BWD=65 DISO_CARVEOUT_BWD=72
python bench.py --dtype f64 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('f64 step', round(d['ms_per_step'],3), d['kernels']['mc_backward'])"
