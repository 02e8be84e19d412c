#!/bin/bash
# round 2, call D: new inference / checkpoint tests + optimizer tests, then compute-sanitizer passes over the 64x64 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_infer.py tests/test_gpu_optim.py tests/test_gpu_smallops.py -m gpu -q -s > gpurun_out/tests_infer.log 2>&1
grep -E "^\[|passed|failed|FAILED|^E  " gpurun_out/tests_infer.log | cut -c1-1200 | tail -40
for tool in memcheck racecheck; do
  ACLGAN_INFER_GRAPHS=0 timeout 700 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py bf16 > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_step ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
