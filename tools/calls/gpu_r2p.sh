#!/bin/bash
# ncu --set full of the window / first-layer kernels (stand-alone launches of tools/bench_layers.py)
mkdir -p gpurun_out /tmp/ncu
for only in "enc 7x7" "final 7x7" "D0 4x4s2 6->64"; do
  tag=$(echo "$only" | tr ' >-' '___')
  ONLY="$only" timeout 600 ncu --set full --clock-control none -k regex:'igemm|wgrad' -s 6 -c 9 -f -o /tmp/ncu/$tag python tools/bench_layers.py 8 > gpurun_out/ncu_$tag.log 2>&1
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > gpurun_out/ncu_$tag.csv 2>/dev/null
done
python - <<'PY'
import csv, glob
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
for f in sorted(glob.glob("gpurun_out/ncu_*.csv")):
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        print(f, "empty"); continue
    hdr = rows[0]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    print("==", f)
    seen = set()
    for r in rows[2:]:
        key = (r[hdr.index("Kernel Name")], r[hdr.index("launch__grid_size")])
        if key in seen: continue
        seen.add(key)
        print({w.replace("smsp__warp_issue_stalled_", "stall_").replace("_per_warp_active.pct", "")[:40]: r[i][:28] for w, i in idx})
PY
