#!/bin/bash
# Round evidence in one GPU call: ncu --set full of the step's kernels, the ncu launch list of bench.py, and the bench line.
tag=${1:-r01}
ncu --set full --clock-control none --import-source on -k regex:"geometry|contract|gather_finalize" -s 5 -c 5 -o gpurun_out/${tag}_full -f python tools/run_one.py 2 None 2 > gpurun_out/${tag}_full.log 2>&1
python tools/summarize_ncu.py gpurun_out/${tag}_full.ncu-rep > gpurun_out/${tag}_ncu_full.json && cp gpurun_out/${tag}_ncu_full.json profiles/${tag}_ncu_full.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 3000 gpurun_out/${tag}_bench.json
