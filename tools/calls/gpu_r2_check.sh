#!/bin/bash
# same-warp barrier init + TMEM allocation: kernel / block / step tests and the racecheck run that reported the pair-kernel prologue
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_blocks.py tests/test_gpu_step.py tests/test_gpu_step256.py tests/test_gpu_optim.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/tests_check.log
SAN_DIM=32 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_step.py bf16 2>&1 | grep -v "Host Frame\|^=========         in \|^$" | awk '!seen[$0]++' | head -30 > gpurun_out/check_racecheck_dim32.log; tail -3 gpurun_out/check_racecheck_dim32.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar 2>/dev/null | cut -c1-200
