#!/bin/bash
# round 2, call y: failing statistical update-parity case in detail, fused finalize (+apply) launches, racecheck report
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py -m gpu -q -x -s -k "two_iterations and p0nf" 2>&1 | grep -E "^\[|assert|Error|iteration|passed|failed" | cut -c1-1500 | tail -12 | tee gpurun_out/tests_y.log
ACLGAN_FOLD_DGRAD=0 timeout 600 python -m pytest tests/test_gpu_optim.py -m gpu -q -x -s -k "two_iterations and p0nf" 2>&1 | grep -E "assert|passed|failed" | cut -c1-600 | tail -3 | tee -a gpurun_out/tests_y.log
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_step.py tests/test_gpu_step256.py tests/test_gpu_infer.py -m gpu -q -x 2>&1 | tail -3 | tee -a gpurun_out/tests_y.log
SAN_DIM=32 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_step.py bf16 2>&1 | grep -v "^=========     at\|^=========         in\|Host Frame\|^$" | head -60 > gpurun_out/racecheck_y.log; head -40 gpurun_out/racecheck_y.log
for v in "X=1" "ACLGAN_FUSE_FINALIZE=0" "X=2"; do
  echo "== $v"; env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_y.err | tee "gpurun_out/bench_y_$v.json" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d.get('e2e', {}).get('value'))"
done
tail -3 gpurun_out/bench_y.err
