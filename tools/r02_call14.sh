#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c_harness.py -m gpu -q > gpurun_out/r02o_t1.log 2>&1; tail -6 gpurun_out/r02o_t1.log
timeout 300 python -m pytest tests/test_reference_vectors.py tests/test_geo_innermodel.py -m gpu -q > gpurun_out/r02o_t2.log 2>&1; tail -3 gpurun_out/r02o_t2.log
python tools/run_one.py 2 None 6
python tools/run_one.py 3 None 6
python tools/run_one.py 1 None 6
timeout 300 python tools/slab_bench.py --scale 0.5 --steps 3 > gpurun_out/r02o_slab05.json 2>gpurun_out/r02o_slab05.err; python -c "
import json; b=json.load(open('gpurun_out/r02o_slab05.json')); print('config5x0.5', b['ms_per_assembly_max_over_ranks'], b['stats_rank0'])"
timeout 300 python tools/sweep_bench.py > gpurun_out/r02o_sweep.json 2> gpurun_out/r02o_sweep.err; tail -c 500 gpurun_out/r02o_sweep.json
