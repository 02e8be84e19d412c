"""perf triage of the tensor-core kernels with DEVICE-side timing (one C call launches the plan N times):
full / MMA-only / TMA-only igemm variants, 128- vs 256-pixel work items, and the wgrad kernel, on the dominant shapes."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch
import aclgan_native as N
import engine as E

eng = E.Engine("bf16")
L = N.lib()
REP = 30

def timed(fn):
    fn(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(REP); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REP * 1e3

def run(cin, cout, k, stride, pad, n, h, label, up_in=False):
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
    plan = N.IgemmPlan(); xs = x.struct()
    N.check(L.aclgan_plan_conv_fwd(C.byref(layer.desc), C.byref(xs), layer.wptr(0), C.byref(o), C.byref(plan)), "plan")
    res = []
    for msub in ("1", "2", "pair"):
        for dbg in ("0", "4"):
            os.environ["ACLGAN_IGEMM_PAIR"] = "1" if msub == "pair" else "0"
            os.environ["ACLGAN_IGEMM_MSUB"] = "1" if msub == "pair" else msub
            os.environ["ACLGAN_IGEMM_DEBUG"] = dbg
            us = timed(lambda r: N.check(L.aclgan_igemm_launch_repeat(C.byref(plan), r, E._sp()), "launch"))
            res.append("m%s/%s %.1fus (%.0fTF)" % (msub, {"0": "full", "1": "mma", "2": "tma", "3": "nostore", "4": "noepi"}[dbg], us, flops / us / 1e6))
    os.environ["ACLGAN_IGEMM_DEBUG"] = "0"
    dy = E.ActT(eng, n, ho, ho, cout, eng.dy_pad(layer), zero=True); dy.buf.normal_()
    wp = N.WgradPlan(); dys = dy.struct()
    N.check(L.aclgan_plan_conv_wgrad(C.byref(layer.desc), C.byref(dys), C.byref(xs), layer.dw().data_ptr(), C.byref(wp)), "wplan")
    for dbg, nm in (("0", "staged"), ("2", "direct"), ("1", "noatomics")):
        os.environ["ACLGAN_WGRAD_DEBUG"] = dbg
        us = timed(lambda r: N.check(L.aclgan_wgrad_launch_repeat(C.byref(wp), r, E._sp()), "wl"))
        res.append("wgrad/%s %.1fus (%.0fTF)" % (nm, us, flops / us / 1e6))
    os.environ["ACLGAN_WGRAD_DEBUG"] = "0"
    res.append("ksplit %d grid %d" % (wp.ksplit, wp.num_taps * wp.m_tiles * wp.n_tiles * wp.ksplit))
    print(label, " | ".join(res), flush=True)

a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(30): a @ a
torch.cuda.synchronize()
run(256, 256, 3, 1, 1, 8, 64, "3x3 256->256 64x64:")
run(256, 128, 5, 1, 2, 8, 128, "5x5 256->128 @128 :")
run(128, 64, 5, 1, 2, 8, 256, "5x5 128->64 @256  :")
run(64, 128, 4, 2, 1, 8, 256, "4x4s2 64->128 @256:")
run(256, 512, 4, 2, 1, 8, 32, "4x4s2 256->512 @32:")
