#!/bin/bash
# round 2, call t: two epilogue warp groups in the segment kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_blocks.py tests/test_gpu_step256.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/tests_t.log
for v in "X=1" "ACLGAN_EPI_GROUPS=1"; do
echo "== $v"
for only in "res 3x3" "up2 main" "up1 main"; do env $v VARIANTS=1 ONLY="$only" python tools/bench_layers.py 8 2>&1 | grep "^| [a-zA-Z]" | grep -v "^| layer" ; done
done | tee gpurun_out/layers_t.txt
for v in "X=1" "ACLGAN_EPI_GROUPS=1" "X=2" "ACLGAN_EPI_GROUPS=1"; do
  echo "== $v"; env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_t.err | tee gpurun_out/bench_t_$v.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, d.get('e2e', {}).get('value'), d.get('roofline', {}).get('achieved'), d.get('roofline', {}).get('frac'))"
done
tail -3 gpurun_out/bench_t.err
