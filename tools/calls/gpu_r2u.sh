#!/bin/bash
# round 2, call u: epilogue groups in all three implicit-GEMM kernels (split / alternating), ncu of the Adam kernel
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_blocks.py tests/test_gpu_step.py tests/test_gpu_step256.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/tests_u.log
for v in "X=1" "ACLGAN_EPI_GROUPS=1"; do
echo "== $v"
env $v VARIANTS=1 python tools/bench_layers.py 8 2>&1 | grep "^| [a-zA-Z]" | grep -v "^| layer"
done | tee gpurun_out/layers_u.txt
for v in "X=1" "ACLGAN_EPI_GROUPS=1" "X=2" "ACLGAN_WINDOW_VSEG=0"; do
  echo "== $v"; env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_u.err | tee gpurun_out/bench_u_$v.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, d.get('e2e', {}).get('value'), d.get('roofline', {}).get('achieved'), d.get('roofline', {}).get('frac'))"
done
tail -3 gpurun_out/bench_u.err
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"adam_kernel" -c 2 -f -o /tmp/ncu/adam python bench.py --profile-step --no-graphs --no-cpu-baseline > gpurun_out/ncu_adam.log 2>&1
ncu -i /tmp/ncu/adam.ncu-rep --page raw --csv > gpurun_out/r2_adam_raw.csv 2>/dev/null
ncu -i /tmp/ncu/adam.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Theoretical Occ|Achieved Occ|Registers|Stall|stall|L2 Hit|DRAM|Issue|Eligible|No Eligible" | head -60 > gpurun_out/ncu_adam_details.txt
