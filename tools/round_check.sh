#!/bin/bash
# on the GPU box: all -m gpu tests (no -x), then bench lines under the given env variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error" | cut -c1-900 | tail -40 | tee gpurun_out/tests.log
for v in "$@"; do
  echo "== bench with $v"
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$v.err | tee gpurun_out/bench_$v.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d.get('e2e', {}).get('value'), d.get('roofline', {}).get('achieved'))"
  tail -3 gpurun_out/bench_$v.err
done
