"""Device-side timing of the weight-gradient kernels on the hot stride-1 shapes, plain vs segment mode."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch  # noqa: E402
import aclgan_native as N  # noqa: E402
import engine as E  # noqa: E402

eng = E.Engine("bf16")
L = N.lib()
SHAPES = [(256, 256, 3, 1, 8, 64), (256, 128, 5, 2, 8, 128), (128, 64, 5, 2, 8, 256)]
for cin, cout, k, pad, n, h in SHAPES:
    w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device="cuda") * 0.02)
    b = torch.nn.Parameter(torch.zeros(cout, device="cuda"))
    arena = E.GradArena(eng.device)
    layer = E.ConvLayer(eng, arena, w, b, 1, pad)
    arena.finalize()
    x = E.ActT(eng, n, h, h, cin, pad, zero=True)
    x.buf.normal_()
    dy = E.ActT(eng, n, h, h, cout, k - 1, zero=True)
    dy.buf.normal_()
    flops = 2.0 * n * h * h * cin * cout * k * k
    print("== wgrad %dx%d %d->%d, %d x %dx%d  (%.1f GFLOP)" % (k, k, cin, cout, n, h, h, flops / 1e9))
    for name, env in (("plain (one CTA per tap)", "0"), ("segment (one CTA per filter row)", "1")):
        os.environ["ACLGAN_WGRAD_SEG"] = env
        plan = N.WgradPlan()
        dys, xs = dy.struct(), x.struct()
        N.check(L.aclgan_plan_conv_wgrad(C.byref(layer.desc), C.byref(dys), C.byref(xs), layer.dw().data_ptr(), C.byref(plan)), "plan")
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        N.check(L.aclgan_wgrad_launch_repeat(C.byref(plan), 2, sp), "warm")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(L.aclgan_wgrad_launch_repeat(C.byref(plan), 10, sp), "launch")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("   %-36s %7.1f us %7.1f TF/s  (grid %d, ksplit %d, tiles m%d n%d)" % (
            name, ms * 1e3, flops / ms / 1e9, plan.num_taps * plan.m_tiles * plan.n_tiles * plan.ksplit, plan.ksplit,
            plan.m_tiles, plan.n_tiles))
