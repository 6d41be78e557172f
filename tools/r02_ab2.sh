#!/bin/bash
# timing-only A/B (no parity run): tools/r02_ab2.sh TAG lib1.so ...
tag=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  n=$(basename $lib .so)
  MOVFEM_B200_LIB=$PWD/$lib timeout 200 python tools/slab_bench.py --scale 0.5 --steps 4 > gpurun_out/${tag}_${n}.json 2>gpurun_out/${tag}_${n}.err
  python -c "
import json; b=json.load(open('gpurun_out/${tag}_${n}.json')); s=b['stats_rank0']; print('$n', round(b['ms_per_assembly_max_over_ranks'],3), 'fused', round(s['ms_fused'],3), 'gather', round(s['ms_gather'],3), 'exact', round(s['ms_exact'],3), 'nflag', s['nflagged'])" || tail -3 gpurun_out/${tag}_${n}.err
done
