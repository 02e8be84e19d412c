#!/bin/bash
# run on the GPU box (via tools/gpu.sh): bench line, ncu launch list of one step-pair, full capture of the top kernels
mkdir -p gpurun_out /tmp/ncu
python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-step --no-graphs --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'igemm_seg_pair_kernel|wgrad_seg_kernel' -s 4 -c 3 -f \
    -o gpurun_out/top_kernels python tools/prof_kernels.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/top_kernels.ncu-rep --page raw --csv > gpurun_out/top_kernels_raw.csv 2>/dev/null
ls -la gpurun_out
