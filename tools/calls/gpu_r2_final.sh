#!/bin/bash
# round 2 final validation on one B200: every -m gpu test, smoke, the bench line (both arms' cheap legs), the ncu launch list of
# one step-pair, a full capture of the top kernels, the in-graph trace, compute-sanitizer memcheck / racecheck
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 3000 gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/final_launches.csv python bench.py --profile-step --no-graphs --no-cpu-baseline > gpurun_out/final_ncu_list.log 2>&1
tail -1 gpurun_out/final_ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'igemm_seg_pair_kernel|wgrad_seg_kernel' -s 4 -c 3 -f \
    -o /tmp/ncu/top_kernels python tools/prof_kernels.py > gpurun_out/final_ncu_full.log 2>&1
ncu -i /tmp/ncu/top_kernels.ncu-rep --page raw --csv > gpurun_out/final_top_kernels_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:"adam_kernel" -c 2 -f -o /tmp/ncu/adam python bench.py --profile-step --no-graphs --no-cpu-baseline > gpurun_out/final_ncu_adam.log 2>&1
ncu -i /tmp/ncu/adam.ncu-rep --page raw --csv > gpurun_out/final_adam_raw.csv 2>/dev/null
python tools/trace_step.py > gpurun_out/final_trace.txt 2>&1; sed -n 3,10p gpurun_out/final_trace.txt
for tool in memcheck racecheck; do
  SAN_DIM=32 timeout 900 compute-sanitizer --tool $tool python tools/sanitize_step.py bf16 2>&1 | grep -v "Host Frame\|^=========         in \|^$" | awk '!seen[$0]++' | head -40 > gpurun_out/final_sanitize_$tool.log; tail -2 gpurun_out/final_sanitize_$tool.log
done
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_step.py bf16 2>&1 | tail -2 > gpurun_out/final_sanitize_racecheck_dim16.log; tail -1 gpurun_out/final_sanitize_racecheck_dim16.log
