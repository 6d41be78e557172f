#!/bin/bash
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02k_smoke.log 2>&1; tail -2 gpurun_out/r02k_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c_harness.py -m gpu -q > gpurun_out/r02k_t1.log 2>&1; tail -6 gpurun_out/r02k_t1.log
timeout 300 python -m pytest tests/test_reference_vectors.py tests/test_geo_innermodel.py -m gpu -q > gpurun_out/r02k_t2.log 2>&1; tail -3 gpurun_out/r02k_t2.log
python tools/run_one.py 2 None 6
python tools/run_one.py 3 None 6
