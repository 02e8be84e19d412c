#!/bin/bash
# round 2, call L: P2 line (selfie2anime fp32x3), sanitizer passes over the final code, ncu --set full of the roofline kernel
mkdir -p gpurun_out /tmp/ncu
python bench.py --steps 10 --warmup 3 --config selfie2anime.yaml --precision fp32x3 --no-cpu-baseline --no-library-bar > gpurun_out/bench_p2.json 2> gpurun_out/bench_p2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_p2.json').read().strip().splitlines()[-1]); print('P2', d['value'], d['ms_per_step'], d['dtype'], d['e2e']['losses'])"
for tool in memcheck racecheck; do
  ACLGAN_INFER_GRAPHS=0 timeout 700 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py bf16 > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_step ok" gpurun_out/sanitize_$tool.log | head -4
done
ACLGAN_INFER_GRAPHS=0 timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step.py fp32x3 > gpurun_out/sanitize_memcheck_fp32x3.log 2>&1
grep -E "ERROR SUMMARY|sanitize_step ok" gpurun_out/sanitize_memcheck_fp32x3.log | head -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'igemm_seg_pair_kernel|wgrad_seg_kernel' -s 4 -c 3 -f \
    -o /tmp/ncu/top_kernels python tools/prof_kernels.py > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/ncu/top_kernels.ncu-rep --page raw --csv > gpurun_out/r2_top_kernels_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_top_kernels_raw.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct"]
idx = [hdr.index(w) for w in want if w in hdr]
print([hdr[i] for i in idx]); print([rows[1][i] for i in idx])
for r in rows[2:]:
    print([r[i][:40] for i in idx])
PY
