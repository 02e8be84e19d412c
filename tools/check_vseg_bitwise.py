"""Is the vertical-segment plan of the first 7x7 conv (pixel windows) bit-identical to the box-per-tap plan?  Same MMAs in the same
order per output pixel, so the raw outputs must match exactly; the fused statistics are summed over differently shaped tiles and
may differ in the last fp32 bits.  usage (GPU box): python tools/check_vseg_bitwise.py [bf16|fp32x3]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch  # noqa: E402
import aclgan_native as N  # noqa: E402
import engine as E  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32x3"
eng = E.Engine(prec)
L = N.lib()
torch.manual_seed(0)
n, h, cin, cout, k, pad = 2, 64, 3, 64, 7, 3
w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device="cuda") * 0.05)
b = torch.nn.Parameter(torch.randn(cout, device="cuda") * 0.1)
arena = E.GradArena(eng.device)
layer = E.ConvLayer(eng, arena, w, b, 1, pad, N.WINDOW_IN)
arena.finalize()
x = E.ActT(eng, n, h, h, cin, pad, cs=8, zero=True)
x.buf.normal_()
res = {}
for vseg in ("0", "1"):
    os.environ["ACLGAN_WINDOW_VSEG"] = vseg
    y = eng.new_dense(n, h, h, cout)
    y.zero_()
    sums = torch.zeros((n, cout, 2), dtype=torch.float64, device="cuda")
    fused = eng.conv_fwd_launch(layer, x, eng._out_dense(y, b), stats=sums)
    torch.cuda.synchronize()
    res[vseg] = (y.clone(), sums.clone(), fused)
y0, s0, f0 = res["0"]
y1, s1, f1 = res["1"]
print("precision %s: raw conv output max |vseg - box| = %.3e (dtype %s, %s), fused stats %s / %s, statistics max rel diff %.3e" % (
    prec, float((y1.double() - y0.double()).abs().max()), y0.dtype, "BIT-IDENTICAL" if torch.equal(y0, y1) else "different",
    f0, f1, float(((s1 - s0).abs() / s0.abs().clamp_min(1e-30)).max())))
