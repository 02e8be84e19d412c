#!/bin/bash
# in-graph timeline + serialised ncu launch list of one step-pair
mkdir -p gpurun_out
python tools/trace_step.py > gpurun_out/trace.txt 2>&1; head -40 gpurun_out/trace.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-step --no-graphs --no-cpu-baseline --no-library-bar > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
python tools/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -45
