#!/bin/bash
# Round evidence in one GPU call: all GPU tests, smoke, the bench line, the ncu launch list of the bench command and ncu --set full
# captures of the step's kernels on config 5 (headline) and config 2.  Summaries are made afterwards with tools/summarize_ncu.py.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${tag}_gputests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log | cut -c1-300
SECONDS=0
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$? wall=${SECONDS}s"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --headline-only --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fused12|gather_finalize|node_kernel|rhs_kernel" -s 5 -c 6 -o gpurun_out/${tag}_full_config5 -f python tools/run_one.py 5 None 2 > gpurun_out/${tag}_full_config5.log 2>&1; echo "ncu config5 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"geometry|contract|exact|gather_finalize|compact" -s 8 -c 8 -o gpurun_out/${tag}_full_config2 -f python tools/run_one.py 2 None 2 > gpurun_out/${tag}_full_config2.log 2>&1; echo "ncu config2 rc=$?"
python - <<PY
import json
b=json.load(open('gpurun_out/${tag}_bench.json'))
print('HEADLINE', b['config']['workload'], 'ms/step', round(b['ms_per_step'],3), 'value', b['value'], {k: round(v,3) for k,v in b['phases_ms'].items()})
print(' e2e ms', round(b['e2e']['ms_per_step'],1), 'keep', round(b['e2e']['keep_pattern_variant']['ms_per_step'],1))
print(' roofline', {k: b['roofline'][k] for k in ('achieved','peak','frac','ms_kernel')}, 'step', b['roofline_step'])
for k, v in b.get('per_config', {}).items():
    print(k, 'ms', round(v['ms_per_step'],4), {a: round(x,4) for a,x in v['phases_ms'].items()}, 'e2e', round(v['e2e']['ms_per_step'],2), round(v['e2e'].get('keep_pattern_ms_per_step',0),2), v['e2e'].get('pageable_ms_per_step'), 'roof', round(v['roofline']['step']['frac_fp64'],3), round(v['roofline']['step']['frac_hbm'],3), 'cpu', v.get('cpu_baseline', {}).get('value'))
print('sweep', {k: b['sweep'][k] for k in ('ms_sweep_max_over_ranks','rank0_cold_frequency_ms','rank0_cached_frequency_ms')})
print('cpu', b['cpu_baseline']['value'], 'clocks', b['clocks'])
PY
