#!/bin/bash
# round-end evidence: tests, both bench arms, ncu launch list of the bench command, one full ncu capture
# of the dominant kernel at 512^3 (DRAM traffic), smoke
TAG=${1:-r1_final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_512.json 2> gpurun_out/bench_ref_512.err; tail -n 3 gpurun_out/bench_ref_512.err
timeout 900 python bench.py > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; tail -n 3 gpurun_out/bench_512.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/launches_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mc_backward -s 1 -c 1 -f -o gpurun_out/prof_bwd512_$TAG python tools/profile_step.py --size 512 --steps 2 --alg mc > gpurun_out/prof_bwd512_$TAG.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/bench_ref_512.json','gpurun_out/bench_512.json'):
    try:
        d=json.load(open(f))
        print(f, d.get('impl','ours'), 'ms', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'cpu', d.get('cpu_baseline'))
        for k,v in (d.get('kernels') or {}).items(): print('  %-20s %8.3f ms x%.0f share %.3f  %s GB/s'%(k, v['ms'], v['launches_per_step'], v['share_of_step'], v['alg_GBps']))
        print(' ', d.get('roofline'), d.get('step_roofline'), d['clocks'], d.get('ref_cuda'))
    except Exception as e: print(f, 'ERR', e)
PY
