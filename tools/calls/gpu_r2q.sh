#!/bin/bash
# round 2, call q: tiled Adam kernel (tests + bench), ncu --set full rows of the row-structured element-wise kernels, trace
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests/test_gpu_optim.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error" | cut -c1-600 | tail -30 | tee gpurun_out/tests_optim.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_q.err | tee gpurun_out/bench_q.json | cut -c1-400
tail -3 gpurun_out/bench_q.err
python tools/trace_step.py > gpurun_out/trace_q.txt 2>&1; head -30 gpurun_out/trace_q.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"rows_kernel|adam_kernel" -c 120 -f -o /tmp/ncu/rows python bench.py --profile-step --no-graphs --no-cpu-baseline > gpurun_out/ncu_rows.log 2>&1
tail -2 gpurun_out/ncu_rows.log
ncu -i /tmp/ncu/rows.ncu-rep --page raw --csv > gpurun_out/r2_rows_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
