#!/bin/bash
# GPU call: parity of the fused linear kernel + the balanced exact kernel, timings, ncu captures
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02b_smoke.log 2>&1; tail -2 gpurun_out/r02b_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02b_t1.log 2>&1; tail -12 gpurun_out/r02b_t1.log
timeout 300 python -m pytest tests/test_reference_vectors.py tests/test_geo_innermodel.py -m gpu -q > gpurun_out/r02b_t2.log 2>&1; tail -3 gpurun_out/r02b_t2.log
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; python -c "
import json; b=json.load(open('gpurun_out/r02b_bench.json')); print('config2', b['ms_per_step'], b['phases_ms'], 'e2e', b['e2e']['ms_per_step'])"
timeout 300 python tools/slab_bench.py --scale 0.5 --steps 3 > gpurun_out/r02b_slab05.json 2>gpurun_out/r02b_slab05.err; python -c "
import json; b=json.load(open('gpurun_out/r02b_slab05.json')); print('config5x0.5', b['ms_per_assembly_max_over_ranks'], b['stats_rank0'])"
MOVFEM_NO_FUSED12=1 timeout 300 python tools/slab_bench.py --scale 0.5 --steps 3 > gpurun_out/r02b_slab05_nofuse.json 2>gpurun_out/r02b_slab05_nofuse.err; python -c "
import json; b=json.load(open('gpurun_out/r02b_slab05_nofuse.json')); print('config5x0.5 unfused', b['ms_per_assembly_max_over_ranks'], b['stats_rank0'])"
timeout 300 python tools/sweep_bench.py > gpurun_out/r02b_sweep.json 2> gpurun_out/r02b_sweep.err; tail -c 700 gpurun_out/r02b_sweep.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"exact_kernel|gather_finalize|compact" -s 4 -c 4 -o gpurun_out/r02b_exact -f python tools/run_one.py 2 None 2 > gpurun_out/r02b_ncu_exact.log 2>&1; tail -2 gpurun_out/r02b_ncu_exact.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fused12" -s 1 -c 1 -o gpurun_out/r02b_fused12 -f python tools/run_one.py 1 None 2 > gpurun_out/r02b_ncu_fused.log 2>&1; tail -2 gpurun_out/r02b_ncu_fused.log
