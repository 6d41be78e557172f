#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_c_harness.py -m gpu -q -k "small or config1_full or sweep or fang or slab or config5_submesh or c_caller" > gpurun_out/r02d_t1.log 2>&1; tail -4 gpurun_out/r02d_t1.log
timeout 300 python tools/slab_bench.py --scale 0.5 --steps 3 > gpurun_out/r02d_slab05.json 2>gpurun_out/r02d_slab05.err; python -c "
import json; b=json.load(open('gpurun_out/r02d_slab05.json')); print('config5x0.5', b['ms_per_assembly_max_over_ranks'], b['stats_rank0'])"
SECONDS=0
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$? wall=${SECONDS}s"; tail -5 gpurun_out/r02d_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02d_bench.json'))
print('HEADLINE', b['config']['workload'], 'ms/step', b['ms_per_step'], 'value', b['value'], 'phases', b['phases_ms'])
print(' e2e', b['e2e']['ms_per_step'], 'keep', b['e2e']['keep_pattern_variant']['ms_per_step'])
print(' roofline', {k: b['roofline'][k] for k in ('achieved','peak','frac','ms_kernel')}, 'step', b['roofline_step'])
for k, v in b.get('per_config', {}).items():
    print(k, 'ms', v['ms_per_step'], v['phases_ms'], 'e2e ms', v['e2e']['ms_per_step'], v['e2e'].get('keep_pattern_ms_per_step'), v['e2e'].get('pageable_ms_per_step'), 'step roof', v['roofline']['step']['frac_fp64'], v['roofline']['step']['frac_hbm'], 'cpu', v.get('cpu_baseline', {}).get('value'))
print('sweep', {k: b['sweep'][k] for k in ('ms_sweep_max_over_ranks','rank0_cold_frequency_ms','rank0_cached_frequency_ms')})
print('cpu', b.get('cpu_baseline'))
print('clocks', b['clocks'])
PY
