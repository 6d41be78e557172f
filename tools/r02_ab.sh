#!/bin/bash
# A/B of library builds on the half-scale config 5 slab: tools/r02_ab.sh TAG lib1.so lib2.so ...   (results in gpurun_out/TAG_*.json)
tag=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "small or config1_full or sweep or slab or config5_submesh" > gpurun_out/${tag}_t1.log 2>&1; tail -3 gpurun_out/${tag}_t1.log
for lib in "$@"; do
  n=$(basename $lib .so)
  for rep in 1 2; do
  MOVFEM_B200_LIB=$PWD/$lib timeout 300 python tools/slab_bench.py --scale 0.5 --steps 5 > gpurun_out/${tag}_${n}_$rep.json 2>gpurun_out/${tag}_${n}.err
  python -c "
import json; b=json.load(open('gpurun_out/${tag}_${n}_$rep.json')); s=b['stats_rank0']; print('$n', round(b['ms_per_assembly_max_over_ranks'],3), 'fused', round(s['ms_fused'],3), 'gather', round(s['ms_gather'],3), 'node', round(s['ms_node'],3), 'nz', s['nz'])"
  done
done
