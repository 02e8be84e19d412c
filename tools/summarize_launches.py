"""ncu launch list (csv of `--metrics gpu__time_duration.sum`) -> markdown table of kernel shares.
usage: python tools/summarize_launches.py gpurun_out/launches.csv "title" > profiles/rN_launches_vK.md"""
import collections
import csv
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[row["Metric Unit"]]
    k = row["Kernel Name"].split("(")[0][:72]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("# %s\n" % title)
print("| launches | total ms | share | kernel |\n|---:|---:|---:|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(__import__("os").environ.get("TOP", "24"))]:
    print("| %d | %.3f | %.1f%% | `%s` |" % (v[0], v[1], 100 * v[1] / tot, k))
print("\nTotal %d launches, %.1f ms serialised (cold-cache per-launch times: compare SHARES)." % (
    sum(v[0] for v in agg.values()), tot))
