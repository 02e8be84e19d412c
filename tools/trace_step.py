"""Kernel timeline of ONE graph-replayed step-pair through torch.profiler (CUPTI): real in-graph kernel durations and
the idle gaps between kernels (what ncu's serialised cold-cache launch list cannot show).
usage (GPU box): python tools/trace_step.py > gpurun_out/trace.txt"""
import collections
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch  # noqa: E402
import yaml  # noqa: E402
import trainer as T  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

cfg = yaml.safe_load(open(os.path.join(ROOT, "acl-gan_b200", "configs", "male2female.yaml")))
cfg["precision"] = "bf16"
torch.manual_seed(0)
tr = T.aclgan_Trainer(cfg).cuda()
b = int(os.environ.get("BATCH", "8"))
xa = torch.rand(b, 3, 256, 256, device="cuda") * 2 - 1
xb = torch.rand(b, 3, 256, 256, device="cuda") * 2 - 1
for _ in range(3):
    tr.dis_update(xa, xb, cfg)
    tr.gen_update(xa, xb, cfg)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.dis_update(xa, xb, cfg)
    tr.gen_update(xa, xb, cfg)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
agg = collections.defaultdict(lambda: [0, 0.0])
busy, gaps, last_end = 0.0, 0.0, None
gap_after = collections.defaultdict(float)
for s, e, name in ks:
    k = name.split("(")[0][:60]
    agg[k][0] += 1
    agg[k][1] += e - s
    if last_end is not None and s > last_end:
        gaps += s - last_end
        gap_after[prev] += s - last_end
    busy += e - s
    if last_end is None or e > last_end:
        last_end, prev = e, k
span = ks[-1][1] - ks[0][0]
print("kernels %d  span %.2f ms  sum of kernel time %.2f ms  idle gaps %.2f ms" % (len(ks), span / 1e3, busy / 1e3, gaps / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%6d %9.3f ms %5.1f%%  avg %7.1f us  gap-after %7.3f ms  %s" % (v[0], v[1] / 1e3, 100 * v[1] / span, v[1] / v[0], gap_after[k] / 1e3, k))
# duration histogram of the heavy kernels: where the launch-bound tail is
print("-- duration buckets (count / total ms) per kernel")
edges = [0, 10, 20, 40, 80, 160, 320, 1e9]
for key in ("igemm_seg_pair", "igemm_seg_kernel", "igemm_pair_kernel", "igemm_kernel", "wgrad_seg", "wgrad_kernel", "block_bwd_apply",
            "block_bwd_reduce", "norm_apply"):
    sel = [e - s0 for s0, e, name in ks if key in name]
    if not sel:
        continue
    cells = []
    for lo, hi in zip(edges[:-1], edges[1:]):
        b = [d for d in sel if lo <= d < hi]
        cells.append("%3d/%5.2f" % (len(b), sum(b) / 1e3))
    print("%-18s %s" % (key, "  ".join(cells)))
print("   buckets (us): " + "  ".join("[%g,%g)" % (lo, hi) for lo, hi in zip(edges[:-1], edges[1:])))
# the longest individual kernels
print("-- longest launches")
for s, e, name in sorted(ks, key=lambda t: t[0] - t[1])[:25]:
    print("%9.1f us  %s" % (e - s, name[:100]))
# concurrency profile: how much of the span runs 0 / 1 / 2 / ... kernels at once, and what runs ALONE (the serial sections)
evts = sorted([(s, 1, name) for s, e, name in ks] + [(e, -1, name) for s, e, name in ks])
depth, last_t, hist = 0, evts[0][0], collections.defaultdict(float)
alone = collections.defaultdict(float)
active = {}
for t, d, name in evts:
    hist[depth] += t - last_t
    if depth == 1 and active:
        alone[next(iter(active)).split("(")[0][:48]] += t - last_t
    last_t = t
    depth += d
    if d > 0:
        active[name] = active.get(name, 0) + 1
    else:
        active[name] -= 1
        if active[name] == 0:
            del active[name]
print("-- concurrency (ms of the span with k kernels in flight): " + "  ".join("%d: %.2f" % (k, v / 1e3) for k, v in sorted(hist.items())))
print("-- kernels that run ALONE (ms): " + "  ".join("%s %.2f" % (k.replace("aclgan::", "").replace("void ", ""), v / 1e3)
                                                    for k, v in sorted(alone.items(), key=lambda kv: -kv[1])[:12]))
