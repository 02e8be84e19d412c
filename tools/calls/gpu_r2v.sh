#!/bin/bash
# round 2, call v: Adam v3 (table-driven 32-bit indexing), 128-register variant of the segment kernels (co-residency)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_optim.py -m gpu -q -k adam_kernel_isolated 2>&1 | tail -3 | tee gpurun_out/tests_v.log
timeout 600 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_blocks.py tests/test_gpu_step256.py -m gpu -q -x 2>&1 | tail -3 | tee -a gpurun_out/tests_v.log
for v in "X=1" "ACLGAN_LIB=/root/repo/acl-gan_b200/libaclgan_b200_r128.so" "ACLGAN_LIB=/root/repo/acl-gan_b200/libaclgan_b200_prev.so" "X=2"; do
  echo "== $v"; env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_v.err | tee "gpurun_out/bench_v_${v##*/}.json" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, d.get('e2e', {}).get('value'), d.get('roofline', {}).get('achieved'), d.get('roofline', {}).get('frac'))"
done
tail -3 gpurun_out/bench_v.err
python tools/trace_step.py > gpurun_out/trace_v.txt 2>&1; sed -n 3,14p gpurun_out/trace_v.txt
