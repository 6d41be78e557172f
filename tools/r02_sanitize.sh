#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py > gpurun_out/r02_san_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|fused12 pipeline" gpurun_out/r02_san_$tool.log | tail -4
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused12_pipeline or unequal" 2>&1 | tail -2
