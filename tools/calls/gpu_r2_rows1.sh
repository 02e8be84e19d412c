#!/bin/bash
# CTAs per SM target of the row kernels: 1 vs 2
mkdir -p gpurun_out
for v in "X=1" "ACLGAN_ROWS_PER_SM=1" "X=2" "ACLGAN_ROWS_PER_SM=1"; do
  echo "== $v"; env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, d.get('e2e', {}).get('value'))"
done
