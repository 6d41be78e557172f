#!/bin/bash
# full GPU test suite + default bench: tools/r02_full.sh TAG
tag=$1
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$? ${SECONDS}s"; tail -4 gpurun_out/${tag}_tests.log
SECONDS=0
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/${tag}_bench.json'))
print('HEADLINE', b['config']['workload'], 'ms/step', round(b['ms_per_step'],2), 'value', round(b['value']/1e6,1), 'M/s phases', {k: round(v,2) for k,v in b['phases_ms'].items()})
print(' e2e', round(b['e2e']['ms_per_step'],1), 'keep', round(b['e2e']['keep_pattern_variant']['ms_per_step'],1))
print(' roofline', {k: b['roofline'][k] for k in ('achieved','peak','frac','ms_kernel')}, 'step', b['roofline_step'])
for k, v in b.get('per_config', {}).items():
    print(k, 'ms', round(v['ms_per_step'],3), {a: round(c,3) for a,c in v['phases_ms'].items()}, 'e2e ms', round(v['e2e']['ms_per_step'],2))
print('sweep', {k: b['sweep'][k] for k in ('ms_sweep_max_over_ranks','rank0_cold_frequency_ms','rank0_cached_frequency_ms')})
print('clocks', b['clocks'])
PY
