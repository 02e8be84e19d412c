"""Stand-alone device timing (warm, back-to-back launches, CUDA events) of forward / data-gradient / weight-gradient kernels
for every convolution shape of the 256x256 networks: where each layer stands against the tensor-core roofline.
usage (GPU box): python tools/bench_layers.py [batch] > gpurun_out/layers.txt"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch  # noqa: E402
import aclgan_native as N  # noqa: E402
import engine as E  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = E.Engine("bf16")
L = N.lib()
SP = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
# name, cin, cout, k, stride, pad, window, n, h (input)
LAYERS = [
    ("enc 7x7 3->64 (window)", 3, 64, 7, 1, 3, N.WINDOW_IN, B, 256),
    ("enc 4x4s2 64->128", 64, 128, 4, 2, 1, 0, B, 256),
    ("enc 4x4s2 128->256", 128, 256, 4, 2, 1, 0, B, 128),
    ("res 3x3 256->256", 256, 256, 3, 1, 1, 0, B, 64),
    ("style 4x4s2 256->256 @64", 256, 256, 4, 2, 1, 0, B, 64),
    ("style 4x4s2 256->256 @32", 256, 256, 4, 2, 1, 0, B, 32),
    ("up1 main 3x3 256->4x128", 256, 512, 3, 1, 1, 0, B, 64),
    ("up2 main 3x3 128->4x64", 128, 256, 3, 1, 1, 0, B, 128),
    ("up1 5x5 256->128 row strips", 256, 128, 5, 1, 2, 0, 2 * B, (2, 128)),
    ("up2 5x5 128->64 row strips", 128, 64, 5, 1, 2, 0, 2 * B, (2, 256)),
    ("final 7x7 64->4 (fold / window)", 64, 4, 7, 1, 3, N.WINDOW_OUT, B, 256),
    ("D0 4x4s2 6->64 (window)", 6, 64, 4, 2, 1, N.WINDOW_IN, 2 * B, 256),
    ("D0 4x4s2 64->128", 64, 128, 4, 2, 1, 0, 3 * B, 128),
    ("D0 4x4s2 128->256", 128, 256, 4, 2, 1, 0, 3 * B, 64),
    ("D0 4x4s2 256->512", 256, 512, 4, 2, 1, 0, 3 * B, 32),
    ("D1 4x4s2 64->128", 64, 128, 4, 2, 1, 0, 3 * B, 64),
    ("D1 4x4s2 256->512", 256, 512, 4, 2, 1, 0, 3 * B, 16),
    ("D2 4x4s2 64->128", 64, 128, 4, 2, 1, 0, 3 * B, 32),
    ("D2 4x4s2 256->512", 256, 512, 4, 2, 1, 0, 3 * B, 8),
]


def timed(fn, reps=10):
    fn(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(reps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("| layer (batch) | GFLOP | forward us | TFLOP/s | dgrad us | TFLOP/s | wgrad us | TFLOP/s |\n|---|---:|---:|---:|---:|---:|---:|---:|")
ONLY = os.environ.get("ONLY")
for name, cin, cout, k, s, pad, window, n, hw in LAYERS:
    if ONLY and ONLY not in name:
        continue
    h, w_ = hw if isinstance(hw, tuple) else (hw, hw)
    wt = torch.nn.Parameter(torch.randn(cout, cin, k, k, device="cuda") * 0.02)
    b = torch.nn.Parameter(torch.zeros(cout, device="cuda"))
    arena = E.GradArena(eng.device)
    layer = E.ConvLayer(eng, arena, wt, b, s, pad, window)
    arena.finalize()
    cs = (16 if s == 2 else 8) if window == N.WINDOW_IN else None
    x = E.ActT(eng, n, h, w_, cin, pad, cs=cs, zero=True)
    x.buf.normal_()
    ho, wo = eng.conv_out_hw(layer, x)
    flops = 2.0 * n * ho * wo * cin * cout * k * k
    # forward
    if window == N.WINDOW_OUT:
        img = torch.empty((n, cout, ho, wo), dtype=torch.float32, device="cuda")
        o = N.OutSpec()
        o.ptr[0] = img.data_ptr()
        o.kind, o.act, o.slope, o.mirror, o.off = N.OUT_F32, N.ACT_TANH, 0.2, 0, 0
        o.sn, o.sy, o.sx, o.sc = cout * ho * wo, wo, 1, ho * wo
        o.N, o.H, o.W, o.C = n, ho, wo, cout
        o.bias, o.bias_n = b.data_ptr(), cout
    else:
        out = E.ActT(eng, n, ho, wo, cout, 1)
        o = eng._out_plane(out, N.ACT_RELU, b)
    plan = N.IgemmPlan()
    xs = x.struct()
    N.check(L.aclgan_plan_conv_fwd(C.byref(layer.desc), C.byref(xs), layer.wptr(0), C.byref(o), C.byref(plan)), "plan")
    t_f = timed(lambda r: N.check(L.aclgan_igemm_launch_repeat(C.byref(plan), r, SP()), "launch"))
    extra = ""
    if window != N.WINDOW_OUT and os.environ.get("VARIANTS"):
        # the forward as the norm blocks run it: dense raw output (+ fused statistics), vs the padded-plane / halo epilogue above
        yd = eng.new_dense(n, ho, wo, ((cout + 63) // 64) * 64)
        od = eng._out_dense(yd, b)
        pd2 = N.IgemmPlan()
        N.check(L.aclgan_plan_conv_fwd(C.byref(layer.desc), C.byref(xs), layer.wptr(0), C.byref(od), C.byref(pd2)), "plan")
        t_dense = timed(lambda r: N.check(L.aclgan_igemm_launch_repeat(C.byref(pd2), r, SP()), "launch"))
        t_stats = float("nan")
        if L.aclgan_igemm_stats_supported(C.byref(pd2)):
            sums = torch.zeros((n, od.C, 2), dtype=torch.float64, device="cuda")
            pd2.out.stats = sums.data_ptr()
            t_stats = timed(lambda r: N.check(L.aclgan_igemm_launch_repeat(C.byref(pd2), r, SP()), "launch"))
        extra = " fwd dense %.1f us, dense+stats %.1f us |" % (t_dense, t_stats)
        if os.environ.get("PROF"):
            # device-side role timers of the segment kernel (cycles, mean over CTAs) without / with the fused statistics
            L.aclgan_igemm_set_prof.argtypes = [C.c_uint64]
            prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
            for tag, st in (("dense", 0), ("dense+stats", sums.data_ptr() if t_stats == t_stats else 0)):
                if tag != "dense" and st == 0:
                    continue
                pd2.out.stats = st
                prof.zero_()
                L.aclgan_igemm_set_prof(prof.data_ptr())
                N.check(L.aclgan_igemm_launch_repeat(C.byref(pd2), 1, SP()), "launch")
                torch.cuda.synchronize()
                L.aclgan_igemm_set_prof(0)
                pr = prof.view(148, 16).double()
                pr = pr[pr[:, 5] > 0]
                m = pr.mean(0)
                extra += ("\n    [%s] cycles/CTA: producer %.0f (wait %.0f) | mma %.0f (wait operands %.0f, wait acc %.0f) | epilogue %.0f (wait %.0f, tile fn %.0f; "
                          "fast %.0f: tmem wait %.0f, chunks %.0f) ctas %d" % (tag, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], pr.shape[0]))
    # data gradient
    dyp = eng.dy_pad(layer)
    dy = E.ActT(eng, n, ho, wo, cout, dyp, cs=8 if window == N.WINDOW_OUT else None, zero=True)
    dy.buf.normal_()
    hp, wp = x.h + 2 * x.pad, x.w + 2 * x.pad
    gcs = x.c if x.c >= 64 else 16
    g = torch.zeros((n, hp, wp, gcs), dtype=torch.float32 if x.c < 64 else torch.bfloat16, device="cuda")
    og = eng._out_dense(g)
    if s == 2:
        og.sy, og.sx = 2 * wp * gcs, 2 * gcs
        og.H, og.W = hp // 2, wp // 2
    pd = N.IgemmPlan()
    dys = dy.struct()
    N.check(L.aclgan_plan_conv_dgrad(C.byref(layer.desc), C.byref(dys), layer.wptr(1), -1 if s == 2 else 0, C.byref(og), C.byref(pd)), "plan dgrad")
    t_d = timed(lambda r: N.check(L.aclgan_igemm_launch_repeat(C.byref(pd), r, SP()), "launch"))
    # weight gradient
    pw = N.WgradPlan()
    N.check(L.aclgan_plan_conv_wgrad(C.byref(layer.desc), C.byref(dys), C.byref(xs), layer.dw().data_ptr(), C.byref(pw)), "plan wgrad")
    t_w = timed(lambda r: N.check(L.aclgan_wgrad_launch_repeat(C.byref(pw), r, SP()), "launch"))
    print("| %s (%d) | %.1f | %.1f | %.0f | %.1f | %.0f | %.1f | %.0f |" % (
        name, n, flops / 1e9, t_f, flops / t_f / 1e6, t_d, flops / t_d / 1e6, t_w, flops / t_w / 1e6) + extra)
