#!/bin/bash
# one iteration of the kernel-tuning loop: parity tests, the 512^3 bench line, sparse-surface timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_FLAGS} > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err
tail -n 5 gpurun_out/bench_512.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_512.json'))
print('ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ref', (d.get('ref_cuda') or {}).get('ms_per_step'))
for k,v in d['kernels'].items(): print('  %-20s %8.3f ms  share %.3f  %s GB/s'%(k, v['ms'], v['share_of_step'], v['alg_GBps']))
print(d['roofline']['kernel'], round(d['roofline']['frac'],3), 'step', round(d['step_roofline']['frac'],3), d['clocks'])
PY
python tools/kernel_breakdown.py sphere 512 dmc 2>&1 | tail -9
python tools/kernel_breakdown.py sphere 512 mc 2>&1 | tail -7
