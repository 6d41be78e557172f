#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "small or config1_full or sweep or fang or slab or config5_submesh or golden or end_to_end or device_resident or keep_pattern or pageable" > gpurun_out/r02q_t1.log 2>&1; tail -4 gpurun_out/r02q_t1.log
python tools/run_one.py 1 None 6
timeout 300 python tools/slab_bench.py --scale 0.5 --steps 3 > gpurun_out/r02q_slab05.json 2>gpurun_out/r02q_slab05.err; python -c "
import json; b=json.load(open('gpurun_out/r02q_slab05.json')); print('config5x0.5', b['ms_per_assembly_max_over_ranks'], b['stats_rank0'])"
