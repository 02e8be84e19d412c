#!/bin/bash
# ncu --set full capture of the main kernels inside one step-pair (bench.py --profile-step); the report stays on the box
# (too large to merge back), its raw-page csv and a per-kernel summary come back in gpurun_out/
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"${KERNELS:-block_bwd_apply_fast|block_bwd_reduce_fast|norm_apply_kernel|wgrad_kernel|igemm_seg_pair_kernel|igemm_pair_kernel|igemm_kernel|igemm_seg_kernel}" \
  -c ${COUNT:-48} -f -o /tmp/ncu/step python bench.py --profile-step --no-graphs --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
ncu -i /tmp/ncu/step.ncu-rep --page raw --csv > gpurun_out/r1_step_kernels_raw.csv 2>/dev/null
ls -la /tmp/ncu gpurun_out
