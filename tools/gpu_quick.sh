#!/bin/bash
# quick validation: GPU parity tests + one bench line at 512^3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err
tail -n 15 gpurun_out/pytest_gpu.log; tail -n 5 gpurun_out/bench_512.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_512.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ref', d.get('ref_cuda'))
for k,v in d['kernels'].items(): print('  %-20s %8.3f ms  share %.3f  %s GB/s'%(k, v['ms'], v['share_of_step'], v['alg_GBps']))
print(d['roofline']); print(d['step_roofline']); print(d['clocks'])
PY
