#!/bin/bash
# round 2: bitwise check of the vertical-segment plan, then the final validation (tools/calls/gpu_r2_final.sh)
mkdir -p gpurun_out
python tools/check_vseg_bitwise.py fp32x3 2>&1 | tail -1 | tee gpurun_out/vseg_bitwise.log
python tools/check_vseg_bitwise.py bf16 2>&1 | tail -1 | tee -a gpurun_out/vseg_bitwise.log
bash tools/calls/gpu_r2_final.sh
