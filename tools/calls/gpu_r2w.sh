#!/bin/bash
# round 2, call w: shared-memory request of the segment kernels (co-residency is gone with 384 x 168 registers): 176 vs 227 KB
mkdir -p gpurun_out
for v in "X=1" "ACLGAN_SEG_SMEM_KB=227" "X=2" "ACLGAN_SEG_SMEM_KB=200"; do
  echo "== $v"; env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_w.err | tee "gpurun_out/bench_w_$v.json" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, d.get('e2e', {}).get('value'), d.get('roofline', {}).get('achieved'), d.get('roofline', {}).get('frac'))"
done
tail -3 gpurun_out/bench_w.err
for v in "X=1" "ACLGAN_SEG_SMEM_KB=227"; do echo "== $v"; for only in "res 3x3" "up1 main" "up2 main"; do env $v VARIANTS=1 ONLY="$only" python tools/bench_layers.py 8 2>&1 | grep "^| [a-zA-Z]" | grep -v "^| layer"; done; done
