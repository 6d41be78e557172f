#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02j_bench2.json 2> gpurun_out/r02j_bench2.err; echo "bench2 rc=$? wall=${SECONDS}s"; tail -5 gpurun_out/r02j_bench2.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r02j_bench2.json') if l.strip().startswith('{')][-1])
print('N=2 HEADLINE', b['config']['workload'], 'ms/step', b['ms_per_step'], 'value', b['value'], 'phases', b['phases_ms'])
print(' e2e', b['e2e']['ms_per_step'], 'keep', b['e2e']['keep_pattern_variant']['ms_per_step'])
print(' sweep', {k: b['sweep'][k] for k in ('ms_sweep_max_over_ranks','n_gpus','value')})
print(' sharding', b['config']['sharding'][:80])
PY
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02j_ref2.json 2> gpurun_out/r02j_ref2.err; echo "ref2 rc=$? wall=${SECONDS}s"; tail -c 300 gpurun_out/r02j_ref2.json
