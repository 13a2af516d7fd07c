#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_nccl_check.py --size 256 > gpurun_out/slab_nccl_$N.log 2>&1; echo "slab exit $?" >> gpurun_out/slab_nccl_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "bench ref exit $?"
grep -v "^$" gpurun_out/slab_nccl_$N.log | tail -n 6; tail -n 3 gpurun_out/bench_n$N.err; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][0]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches')}, d['e2e'], d['clocks'])
print(open('gpurun_out/bench_ref_n$N.json').read()[:600])
PY
