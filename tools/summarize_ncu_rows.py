"""ncu --set full raw-page csv (ncu -i rep --page raw --csv) -> markdown table of the HBM-bound kernels: duration, DRAM bytes,
achieved DRAM GB/s, instructions, occupancy.  usage: python tools/summarize_ncu_rows.py gpurun_out/r2_rows_raw.csv [peak GB/s]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6550.0
hdr, units = rows[0], rows[1]
col = hdr.index


def f(r, n):
    try:
        return float(r[col(n)].replace(",", ""))
    except ValueError:
        return float("nan")


agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[col("Kernel Name")].split("(")[0]
    name = name[name.rfind("::", 0, name.find("<") if "<" in name[12:] else len(name)) + 2:] if "::" in name else name
    key = (name, r[col("launch__grid_size")], int(f(r, "dram__bytes_read.sum") / 4))
    agg.setdefault(key, []).append(r)
print("| launches | kernel | grid | duration us | DRAM read MB | DRAM write MB | DRAM GB/s | frac of %.0f | warp instr | warps active %% | regs |" % peak)
print("|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for key, rs in agg.items():
    avg = lambda n: sum(f(r, n) for r in rs) / len(rs)
    d = avg("gpu__time_duration.sum")
    bw = (avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum")) / d * 1e3
    print("| %d | `%s` | %s | %.1f | %.1f | %.1f | %.0f | %.2f | %.2e | %.0f | %d |" % (
        len(rs), key[0], key[1], d, avg("dram__bytes_read.sum"), avg("dram__bytes_write.sum"), bw, bw / peak,
        avg("smsp__inst_executed.sum"), avg("sm__warps_active.avg.pct_of_peak_sustained_active"), avg("launch__registers_per_thread")))
