#!/bin/bash
# bench.py at N ranks (x-slab strong scaling of config 5 + frequency-sharded sweep), as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench N=$N rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
b=json.loads([l for l in open('gpurun_out/r02_bench_${N}gpu.json') if l.strip().startswith('{')][-1])
print('N=$N ms/step', b['ms_per_step'], 'value', b['value'], 'phases', {k: round(v,2) for k,v in b['phases_ms'].items()})
print(' e2e ms', b['e2e']['ms_per_step'], 'keep', b['e2e']['keep_pattern_variant']['ms_per_step'])
print(' sweep', {k: b['sweep'][k] for k in ('ms_sweep_max_over_ranks','n_gpus','value')})
PY
free -g | head -2
