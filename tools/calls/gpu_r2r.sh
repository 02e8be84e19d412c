#!/bin/bash
# round 2, call r: Adam v2, vertical window segments (fwd + resident weights, wgrad IN / OUT + merged taps): kernel tests, layer
# timings under the switches, role timers of the stats epilogue, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_blocks.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/tests_r.log
timeout 300 python -m pytest tests/test_gpu_optim.py -m gpu -q -k adam_kernel_isolated 2>&1 | tail -3 | tee -a gpurun_out/tests_r.log
for v in "X=1" "ACLGAN_WINDOW_VSEG=0" "ACLGAN_SEG_BRES=0" "ACLGAN_WGRAD_MERGE=0"; do
  echo "== $v"
  for only in "enc 7x7" "final 7x7"; do env $v ONLY="$only" python tools/bench_layers.py 8 2>&1 | grep "^| [a-zA-Z]" ; done
done | tee gpurun_out/layers_r.txt
for only in "res 3x3" "up2 main" "enc 7x7"; do VARIANTS=1 PROF=1 ONLY="$only" python tools/bench_layers.py 8 2>&1 | grep -v "^| layer\|^|---" ; done | tee gpurun_out/prof_r.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r.err | tee gpurun_out/bench_r.json | cut -c1-300
tail -3 gpurun_out/bench_r.err
python tools/trace_step.py > gpurun_out/trace_r.txt 2>&1; head -12 gpurun_out/trace_r.txt
