#!/bin/bash
# Launch list of one short request (ncu, cold-cache serialised per-launch times: compare SHARES).
# Usage (on the GPU box, via gpurun):  bash tools/profile_launches.sh
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --ddim-steps 2 --no-graph --no-cpu-baseline --no-gpu-reference > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt
tail -40 gpurun_out/launches_summary.txt
