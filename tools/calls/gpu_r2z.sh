#!/bin/bash
# round 2, call z: bisect the marginal update-parity case (dis_2 first conv, 1.3e-2 vs 1e-2), fused finalize with shared coefficients
mkdir -p gpurun_out
for v in "X=1" "ACLGAN_ELEMENTWISE_ROWS=0" "ACLGAN_ROWS_PER_SM=6" "ACLGAN_FUSE_FINALIZE=0" "ACLGAN_WINDOW_VSEG=0" "ACLGAN_EPI_GROUPS=1" "ACLGAN_WGRAD_MERGE=0" "ACLGAN_FOLD=0"; do
  echo "== $v"
  env $v timeout 300 python -m pytest tests/test_gpu_optim.py -m gpu -q -x -s -k "two_iterations and p0nf" 2>&1 | grep -E "iteration 1 dis\]|assert not|passed|failed" | cut -c1-400 | tail -3
done | tee gpurun_out/bisect_z.log
timeout 600 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_step256.py -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/tests_z.log
for v in "X=1" "ACLGAN_FUSE_FINALIZE=0" "X=2" "ACLGAN_FUSE_FINALIZE=0"; do
  echo "== $v"; env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_z.err | tee "gpurun_out/bench_z_$v.json" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d.get('e2e', {}).get('value'))"
done
tail -3 gpurun_out/bench_z.err
