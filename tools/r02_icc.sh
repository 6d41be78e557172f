#!/bin/bash
# instruction-cache metrics of fused12_kernel for several library builds: tools/r02_icc.sh TAG lib1.so ...
tag=$1; shift
mkdir -p gpurun_out
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
for lib in "$@"; do
  n=$(basename $lib .so)
  MOVFEM_B200_LIB=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none -k regex:${KREGEX:-fused12} -c ${NLAUNCH:-1} --csv --log-file gpurun_out/${tag}_${n}.csv ${DRIVER:-python tools/slab_bench.py --scale 0.5 --steps 1} > /dev/null 2>gpurun_out/${tag}_${n}.err
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${tag}_${n}.csv')) if len(r)>5]
h=rows[0]; mi=h.index('Metric Name'); vi=h.index('Metric Value')
ki=h.index('Kernel Name'); ii=h.index('ID')
import collections
per=collections.OrderedDict()
for r in rows[1:]: per.setdefault((r[ii], r[ki][:40]), {})[r[mi].replace('smsp__average_warps_issue_stalled_','st_').replace('_per_issue_active.ratio','').replace('.pct_of_peak_sustained_elapsed','').replace('.avg.pct_of_peak_sustained_active','')] = r[vi]
for k,v in per.items(): print('$n', k, v)
raise SystemExit
print('$n', {r[mi].replace('smsp__average_warps_issue_stalled_','st_').replace('_per_issue_active.ratio','').replace('.pct_of_peak_sustained_elapsed','').replace('.avg.pct_of_peak_sustained_active',''): r[vi] for r in rows[1:]})
PY
done
