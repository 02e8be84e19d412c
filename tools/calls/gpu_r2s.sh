#!/bin/bash
# round 2, call s: register-resident statistics / halo in the direct epilogue; vseg final-conv dgrad
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_blocks.py tests/test_gpu_step.py tests/test_gpu_step256.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/tests_s.log
for only in "res 3x3" "up2 main" "enc 7x7" "final 7x7" "D0 4x4s2 64->128" "enc 4x4s2 64"; do VARIANTS=1 ONLY="$only" python tools/bench_layers.py 8 2>&1 | grep "^| [a-zA-Z]" | grep -v "^| layer" ; done | tee gpurun_out/layers_s.txt
for v in "X=1" "ACLGAN_EPI_DIRECT=0" "X=2"; do
  echo "== $v"; env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_s.err | tee gpurun_out/bench_s_$v.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, d.get('e2e', {}).get('value'), d.get('roofline', {}).get('achieved'), d.get('roofline', {}).get('frac'))"
done
tail -3 gpurun_out/bench_s.err
