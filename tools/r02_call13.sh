#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"exact_kernel" -s 2 -c 2 -o gpurun_out/r02l_exact -f python tools/run_one.py 2 None 2 > gpurun_out/r02l_ncu_exact.log 2>&1; tail -2 gpurun_out/r02l_ncu_exact.log
