#!/bin/bash
# round 2, call x: backward row kernels: raw loads in flight, 4 pixels per thread, CTAs per SM target
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_step256.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/tests_x.log
for v in "X=1" "ACLGAN_LIB=/root/repo/acl-gan_b200/libaclgan_b200_prev.so" "ACLGAN_BWD_UNR=2" "ACLGAN_ROWS_PER_SM=3" "ACLGAN_ROWS_PER_SM=2" "ACLGAN_ROWS_PER_SM=12"; do
  echo "== $v"; env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_x.err | tee "gpurun_out/bench_x_${v##*/}.json" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, d.get('e2e', {}).get('value'))"
done
tail -3 gpurun_out/bench_x.err
python tools/trace_step.py > gpurun_out/trace_x.txt 2>&1; sed -n 3,8p gpurun_out/trace_x.txt
