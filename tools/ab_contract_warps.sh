#!/bin/bash
# A/B of the consumer-warp counts of contract_kernel (round-2 experiment; see tools/micro/README.md, tile_bench).
# Build here (no GPU needed):   tools/ab_contract_warps.sh build
# Measure on the GPU box:       gpurun -- 'tools/ab_contract_warps.sh run'     (config 2 via bench.py, config 1/4 via sweep_bench)
set -e
cd "$(dirname "$0")/.."
variants=("base:" "c12w7:-DMOVFEM_CON12_W=7" "c12w3:-DMOVFEM_CON12_W=3" "c36w11:-DMOVFEM_CON36_W=11" "c36pw11:-DMOVFEM_CON36P_W=11" "c36pw15:-DMOVFEM_CON36P_W=15")
if [ "$1" = build ]; then
  mkdir -p ab
  for v in "${variants[@]}"; do
    name=${v%%:*}; flags=${v#*:}
    [ "$name" = base ] && continue
    make -s -C movfem_b200/csrc OUT=$PWD/ab/lib_$name.so EXTRA="$flags" -B 2>&1 | grep -iE "error" || true
    ls -la ab/lib_$name.so
  done
  exit 0
fi
mkdir -p gpurun_out
for v in "${variants[@]}"; do
  name=${v%%:*}
  if [ "$name" = base ]; then unset MOVFEM_B200_LIB; else export MOVFEM_B200_LIB=$PWD/ab/lib_$name.so; fi
  timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/abw_$name.json 2> gpurun_out/abw_$name.err || true
  python -c "
import json; b=json.load(open('gpurun_out/abw_$name.json')); print('$name', 'config2 ms/step', round(b['ms_per_step'],4), {k: round(x,4) for k,x in b['phases_ms'].items()})" || true
  timeout 200 python tools/sweep_bench.py > gpurun_out/abw_sweep_$name.json 2> gpurun_out/abw_sweep_$name.err || true
  tail -c 600 gpurun_out/abw_sweep_$name.json; echo
done
