#!/bin/bash
# Round-2 A/B sweep: every experiment DESIGN.md section 7 lists that needs no new kernel code.  All variants compute the
# same bits as the default library (warp counts, cache policies and chunk sizes do not change any summation order).
#   tools/ab_round2.sh build        here, no GPU: builds ab/lib_<variant>.so (about 20 s each)
#   gpurun --timeout 900 -- 'tools/ab_round2.sh run'      on the B200: bench.py (config 2) per variant, then the env-var runs
# Results: gpurun_out/ab2_<variant>.json (bench lines) and the one-line summaries on stdout.
# ab/ is git-ignored but travels with every gpurun snapshot (about 5 MB per library): build only what the call needs
# (ONLY="...") and remove ab/ afterwards.  A full run is roughly 30 variants x (hash 10 s + bench 30 s) + the sweeps: split it
# over two or three calls with ONLY.
set -e
cd "$(dirname "$0")/.."
variants=(
  "base:"
  "c12w7:-DMOVFEM_CON12_W=7"            # me=12: 7 consumers + producer = 8 warps per CTA (even over the SMSPs)
  "c12w3:-DMOVFEM_CON12_W=3"            # me=12: 3 + 1 = 4 warps per CTA
  "c36w11:-DMOVFEM_CON36_W=11"          # me=36: 11 + 1 = 12 warps
  "c36pw11:-DMOVFEM_CON36P_W=11"        # me=36 GPML: 11 + 1 = 12 warps (today 12 + 1 = 13)
  "c36pw15:-DMOVFEM_CON36P_W=15"        # me=36 GPML: 15 + 1 = 16 warps
  "c36pw8:-DMOVFEM_CON36P_W=8"          # me=36 GPML: 8 + 1 warps (an item has 6 or 9 tiles and the ring only 2 stages: idle warps run ahead into the barrier)
  "gld1:-DMOVFEM_GATHER_LD=1"           # gather: __ldcs on the K/M reads
  "gld2:-DMOVFEM_GATHER_LD=2"           # gather: ld.global.nc.L1::no_allocate
  "gld3:-DMOVFEM_GATHER_LD=3"           # gather: cp.async.cg 16-byte copies global -> shared (no register round trip, no L1)
  "gst1:-DMOVFEM_GATHER_ST=1"           # gather: streaming stores of A
  "kmst1:-DMOVFEM_KM_ST=1"              # contraction: streaming stores of K_e/M_e
  "hints:-DMOVFEM_GATHER_LD=2 -DMOVFEM_GATHER_ST=1 -DMOVFEM_KM_ST=1"
  "geopf:-DMOVFEM_GEO_PREFETCH=1"       # geometry: L2 prefetch of the next batch's node records one phase before the bulk copies (19 % of the samples wait for them)
  "geoearly:-DMOVFEM_GEO_EARLY_REQ=1"   # geometry: next batch's bulk copies issued after B1 instead of after B2 (B2 reads a small z/x/y copy)
  "geoearly2:-DMOVFEM_GEO_EARLY_REQ=1 -DMOVFEM_RHS_PER_SLOT=1"
  "rhsslot:-DMOVFEM_RHS_PER_SLOT=1"     # geometry: RHS phase with EB*MEP tasks instead of EB*MEP/4 (72 of 256 threads busy today)
  "geoall:-DMOVFEM_RHS_PER_SLOT=1 -DMOVFEM_GEO_PREFETCH=1"
  "tabg:-DMOVFEM_TAB_GLOBAL=1"          # contraction: operand table filled from global memory instead of lane-distinct constant loads (4 % of the samples)
  "u1:-DMOVFEM_CON_UNROLL=1"            # contraction: Gauss-point loop not unrolled (today 3)
  "u9:-DMOVFEM_CON_UNROLL=9"            # ... unrolled by 9
  "st4:-DMOVFEM_CON36_STAGES=4"         # me=36: ring of 4 class blocks (today 5)
  "tall7:-DMOVFEM_TALL_TILES=1 -DMOVFEM_CON36_W=7 -DMOVFEM_CON54_W=7"     # 8x4 contraction tiles, 7 consumers + producer = 8 warps, 236 regs, no spills
  "tall11:-DMOVFEM_TALL_TILES=1 -DMOVFEM_CON36_W=11 -DMOVFEM_CON54_W=11"  # ... 12 warps, 168 regs (the per-SMSP limit at 3 warps), 376 B of spills
  "tallfold:-DMOVFEM_TALL_TILES=2 -DMOVFEM_CON36_W=8 -DMOVFEM_CON54_W=8"  # 8x4 tiles, producer folded into consumer warp 0: 8 even warps, 236 regs, no spills
  "tall3fold:-DMOVFEM_TALL_TILES=2 -DMOVFEM_TALL_RG=3 -DMOVFEM_CON_UNROLL=1 -DMOVFEM_CON36_W=8 -DMOVFEM_CON54_W=8"  # 12x4 tiles: 254 regs, no spills only with the g loop not unrolled
  "fold16:-DMOVFEM_FOLD_PRODUCER=1 -DMOVFEM_CON12_W=8 -DMOVFEM_CON36_W=16 -DMOVFEM_CON36P_W=12 -DMOVFEM_CON54_W=16"   # 4x4 tiles, no producer warp: 16 (8, 12) consumer warps, all SMSPs even; 128 regs with ~150 B of spills
  "fold12:-DMOVFEM_FOLD_PRODUCER=1 -DMOVFEM_CON12_W=4 -DMOVFEM_CON36_W=12 -DMOVFEM_CON36P_W=8 -DMOVFEM_CON54_W=12"    # ... 12 (4, 8) consumer warps, up to 168 regs
  "st3:-DMOVFEM_CON36_STAGES=3"         # ... 3 (6 would need 238 kB > 227 kB)
)
# ONLY="name1 name2 ..." restricts build/run to those variants (base always runs)
if [ -n "$ONLY" ]; then
  keep=("base:")
  for v in "${variants[@]}"; do for o in $ONLY; do [ "${v%%:*}" = "$o" ] && keep+=("$v"); done; done
  variants=("${keep[@]}")
fi
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
summary() { python -c "
import json,sys; b=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step', round(b['ms_per_step'],4), {k: round(x,4) for k,x in b['phases_ms'].items()})" "$1" "$2" || true; }
for v in "${variants[@]}"; do
  name=${v%%:*}
  if [ "$name" = base ]; then unset MOVFEM_B200_LIB; else export MOVFEM_B200_LIB=$PWD/ab/lib_$name.so; [ -f "$MOVFEM_B200_LIB" ] || continue; fi
  timeout 120 python tools/ab_hash.py > gpurun_out/ab2_hash_$name.txt 2> gpurun_out/ab2_hash_$name.err || true
  if [ "$name" != base ] && ! cmp -s gpurun_out/ab2_hash_base.txt gpurun_out/ab2_hash_$name.txt; then echo "$name: RESULTS DIFFER FROM THE DEFAULT LIBRARY"; fi
  timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab2_$name.json 2> gpurun_out/ab2_$name.err || true
  summary gpurun_out/ab2_$name.json $name
done
unset MOVFEM_B200_LIB
[ -n "$SKIP_ENV" ] && exit 0      # SKIP_ENV=1: library variants only
# env-var experiments on the default library: scratch chunks small enough to stay in the 126 MB L2 between geometry_kernel
# and contract_kernel (config 2 writes 124 MB of Q|T for the unstretched list), and the structured gather
for mb in 24 48 96; do
  MOVFEM_SCRATCH_MB=$mb timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab2_scratch$mb.json 2> gpurun_out/ab2_scratch$mb.err || true
  summary gpurun_out/ab2_scratch$mb.json scratch${mb}MB
done
for cfg in "48:64" "32:48" "64:96"; do   # scratch chunk : persisting-L2 carve-out (MB): the chunk stays in L2 between the two kernels
  MOVFEM_SCRATCH_MB=${cfg%%:*} MOVFEM_L2_PERSIST_MB=${cfg#*:} timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab2_l2p_${cfg%%:*}.json 2> gpurun_out/ab2_l2p_${cfg%%:*}.err || true
  summary gpurun_out/ab2_l2p_${cfg%%:*}.json scratch${cfg%%:*}MB+persist${cfg#*:}MB
done
MOVFEM_GATHER_TEMPLATE=1 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab2_tmpl.json 2> gpurun_out/ab2_tmpl.err || true
summary gpurun_out/ab2_tmpl.json gather_template
[ -n "$SKIP_LINEAR" ] && exit 0   # SKIP_LINEAR=1: no config-5 / config-4 runs
# the me=12 variants matter on the linear meshes: config 5 at half scale (cold single-frequency assembly, GPML Fang) ...
for name in base c12w7 c12w3 fold16 fold12 geoearly geoearly2 rhsslot tabg gld3 hints; do
  if [ "$name" = base ]; then unset MOVFEM_B200_LIB; else export MOVFEM_B200_LIB=$PWD/ab/lib_$name.so; [ -f "$MOVFEM_B200_LIB" ] || continue; fi
  timeout 200 python tools/slab_bench.py --scale 0.5 --steps 3 > gpurun_out/ab2_slab_$name.json 2> gpurun_out/ab2_slab_$name.err || true
  echo "config5 x0.5 $name: $(python -c "
import json,sys; b=json.load(open(sys.argv[1])); s=b.get('stats_rank0',{}); print(round(b['ms_per_assembly_max_over_ranks'],3),'ms', {k: round(s[k],3) for k in ('ms_node','ms_geometry','ms_contract','ms_gather') if k in s})" gpurun_out/ab2_slab_$name.json 2>/dev/null)"
done
# ... and the config 4 sweep (cold first frequency + cached ones)
for name in base c12w7 c12w3; do
  if [ "$name" = base ]; then unset MOVFEM_B200_LIB; else export MOVFEM_B200_LIB=$PWD/ab/lib_$name.so; [ -f "$MOVFEM_B200_LIB" ] || continue; fi
  timeout 200 python tools/sweep_bench.py > gpurun_out/ab2_sweep_$name.json 2> gpurun_out/ab2_sweep_$name.err || true
  echo "sweep $name: $(tail -c 500 gpurun_out/ab2_sweep_$name.json)"
done
