#!/bin/bash
# per-kernel times + FP64/LSU utilisation of one cold assembly: tools/kernel_times.sh <config> [dirichlet]
cfg=${1:-2}; dir=${2:-None}
ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"geometry|contract|gather|node_kernel|rhs" -s 6 -c 8 --csv --log-file gpurun_out/kt.csv python tools/run_one.py $cfg $dir 2 > gpurun_out/kt_run.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/kt.csv")) if len(r)>10]
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[0],r[4][:64]),{})[r[-3].split('.')[0].replace('sm__pipe_','').replace('l1tex__data_pipe_','').replace('smsp__','')]=r[-1]
for k,v in d.items(): print(k[1], v)
PY
