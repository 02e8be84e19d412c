"""Launches the dominant tensor-core kernels a few times on realistic shapes (for ncu captures):
    ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 2 -c 2 -o gpurun_out/igemm python tools/prof_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch  # noqa: E402
import aclgan_native as N  # noqa: E402
import engine as E  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
eng = E.Engine(prec)
n, c, h = 8, 256, 64
w = torch.nn.Parameter(torch.randn(c, c, 3, 3, device="cuda") * 0.02)
b = torch.nn.Parameter(torch.zeros(c, device="cuda"))
arena = E.GradArena(eng.device)
layer = E.ConvLayer(eng, arena, w, b, 1, 1)
arena.finalize()
x = E.ActT(eng, n, h, h, c, 1, zero=True)
x.buf.normal_()
out = E.ActT(eng, n, h, h, c, 1)
o = eng._out_plane(out, N.ACT_NONE, b)
dy = E.ActT(eng, n, h, h, c, 2, zero=True)
dy.buf.normal_()
for _ in range(4):
    eng.conv_fwd_launch(layer, x, o)
    eng.conv_wgrad(layer, dy, x)
    g = eng.conv_dgrad(layer, dy, x)
torch.cuda.synchronize()
print("done")
