#!/bin/bash
# One bounded GPU call that refreshes the round's evidence, most important first (each step has its own timeout):
# reference-vector parity, smoke, bench line, the rest of the GPU tests, the ncu launch list.
tag=${1:-r01b}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_reference_vectors.py -m gpu -q > gpurun_out/${tag}_t_ref.log 2>&1; echo "ref-vector tests rc=$?"; tail -3 gpurun_out/${tag}_t_ref.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/${tag}_bench.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout ${2:-420} python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${tag}_t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/${tag}_t_gpu.log
