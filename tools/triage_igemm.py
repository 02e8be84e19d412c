"""perf triage of the igemm kernel: full / MMA-only / TMA-only timings on the dominant conv shapes"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch
import aclgan_native as N
import engine as E

eng = E.Engine("bf16")
def run(cin, cout, k, stride, pad, n, h, label):
    w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device="cuda") * 0.02)
    b = torch.nn.Parameter(torch.zeros(cout, device="cuda"))
    arena = E.GradArena(eng.device)
    layer = E.ConvLayer(eng, arena, w, b, stride, pad)
    arena.finalize()
    x = E.ActT(eng, n, h, h, cin, pad, zero=True); x.buf.normal_()
    ho = (h + 2 * pad - k) // stride + 1
    out = E.ActT(eng, n, ho, ho, cout, 1)
    o = eng._out_plane(out, N.ACT_NONE, b)
    flops = 2.0 * n * ho * ho * cout * cin * k * k
    res = []
    for msub in ("1", "2"):
        for dbg in ("0", "1", "2"):
            os.environ["ACLGAN_IGEMM_MSUB"] = msub
            os.environ["ACLGAN_IGEMM_DEBUG"] = dbg
            for _ in range(20):
                eng.conv_fwd_launch(layer, x, o)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                eng.conv_fwd_launch(layer, x, o)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 50
            res.append("msub%s/%s: %.1fus (%.0f TF)" % (msub, {"0": "full", "1": "mma-only", "2": "tma-only"}[dbg], ms * 1e3, flops / ms / 1e9))
    print(label, " | ".join(res), flush=True)

# spin the clocks up
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(50): a @ a
torch.cuda.synchronize()
run(256, 256, 3, 1, 1, 8, 64, "3x3 256->256 64x64:")
run(256, 128, 5, 1, 2, 8, 128, "5x5 256->128 128x128:")
run(128, 64, 5, 1, 2, 8, 256, "5x5 128->64 256x256:")
run(64, 128, 4, 2, 1, 8, 256, "4x4s2 64->128 256->128:")
run(256, 512, 4, 2, 1, 8, 32, "4x4s2 256->512 32->16:")
