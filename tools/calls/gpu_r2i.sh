#!/bin/bash
# ncu --set full of the element-wise kernels inside one step-pair
mkdir -p gpurun_out /tmp/ncu
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"${KERNELS:-block_bwd_apply_fast|block_bwd_reduce_fast|norm_apply_kernel}" \
  -c ${COUNT:-36} -f -o /tmp/ncu/step python bench.py --profile-step --no-graphs --no-cpu-baseline --no-library-bar > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
ncu -i /tmp/ncu/step.ncu-rep --page raw --csv > gpurun_out/r2_step_kernels_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_step_kernels_raw.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__grid_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "launch__occupancy_limit_registers", "launch__block_size"]
idx = [hdr.index(w) for w in want if w in hdr]
print([hdr[i] for i in idx])
for r in rows[2:]:
    print([r[i][:48] for i in idx])
PY
