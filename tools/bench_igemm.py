"""Device-side timing of the forward implicit GEMM on the hot layer shapes under several kernel configurations
(environment switches are read at plan / launch time).  usage (GPU box): python tools/bench_igemm.py"""
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
SHAPES = [  # cin, cout, k, pad, n, h[, stride]
    (256, 256, 3, 1, 8, 64),
    (256, 128, 5, 2, 8, 128),
    (128, 64, 5, 2, 8, 256),
    (64, 4, 7, 3, 8, 256),
    (64, 128, 4, 1, 8, 256, 2),      # content / style encoder down-sampling, D layers (4x4 stride 2)
    (128, 256, 4, 1, 8, 128, 2),
    (256, 512, 4, 1, 24, 32, 2),
]
if os.environ.get("ONLY_S2"):
    SHAPES = [s for s in SHAPES if len(s) > 6]
if os.environ.get("ONLY7"):
    SHAPES = [s for s in SHAPES if s[2] == 7]
    CONFIGS = CONFIGS[2:3]
CONFIGS = [
    ("box-per-tap plans, plain kernels", dict(ACLGAN_SEG="0")),
    ("segment plans on plain kernels", dict(ACLGAN_SEG="1", ACLGAN_SEGK="0")),
    ("segment kernel", dict(ACLGAN_SEG="1", ACLGAN_SEGK="1")),
    ("segment kernel, direct epilogue (no smem)", dict(ACLGAN_SEG="1", ACLGAN_SEGK="1", ACLGAN_EPI_DIRECT="1")),
    ("segment kernel, one CTA", dict(ACLGAN_SEG="1", ACLGAN_SEGK="1", ACLGAN_IGEMM_PAIR="0")),
    ("segment kernel, aligned windows (wrong results)", dict(ACLGAN_SEG="1", ACLGAN_SEGK="1", ACLGAN_SEG_DEBUG="1")),
    ("plain pair kernel, epilogue without stores", dict(ACLGAN_SEG="1", ACLGAN_SEGK="0", ACLGAN_IGEMM_DEBUG="3")),
    ("plain pair kernel, no epilogue", dict(ACLGAN_SEG="1", ACLGAN_SEGK="0", ACLGAN_IGEMM_DEBUG="4")),
    ("plain pair kernel, generic epilogue", dict(ACLGAN_SEG="1", ACLGAN_SEGK="0", ACLGAN_IGEMM_DEBUG="5")),
    ("segment kernel, MMA only", dict(ACLGAN_SEG="1", ACLGAN_SEGK="1", ACLGAN_IGEMM_DEBUG="1")),
    ("segment kernel, TMA only", dict(ACLGAN_SEG="1", ACLGAN_SEGK="1", ACLGAN_IGEMM_DEBUG="2")),
]
KEYS = sorted({k for _, c in CONFIGS for k in c})
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
L.aclgan_igemm_set_prof.argtypes = [C.c_uint64]
for shp in SHAPES:
    cin, cout, k, pad, n, h = shp[:6]
    stride = shp[6] if len(shp) > 6 else 1
    w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device="cuda") * 0.02)
    b = torch.nn.Parameter(torch.zeros(cout, device="cuda"))
    arena = E.GradArena(eng.device)
    layer = E.ConvLayer(eng, arena, w, b, stride, pad, N.WINDOW_OUT if cout <= 8 else N.WINDOW_NONE)
    arena.finalize()
    x = E.ActT(eng, n, h, h, cin, pad, zero=True)
    x.buf.normal_()
    if cout <= 8:
        img = torch.empty((n, cout, h, h), dtype=torch.float32, device="cuda")
        o = N.OutSpec()
        o.ptr[0] = img.data_ptr()
        o.kind, o.act, o.slope, o.mirror, o.off = N.OUT_F32, N.ACT_TANH, 0.2, 0, 0
        o.sn, o.sy, o.sx, o.sc = cout * h * h, h, 1, h * h
        o.N, o.H, o.W, o.C = n, h, h, cout
        o.bias, o.bias_n = b.data_ptr(), cout
    else:
        ho = (h + 2 * pad - k) // stride + 1
        y = eng.new_dense(n, ho, ho, cout)
        o = eng._out_dense(y, b)
    ho = (h + 2 * pad - k) // stride + 1
    flops = 2.0 * n * ho * ho * cin * cout * k * k
    print("== %dx%d s%d %d->%d, %d x %dx%d  (%.1f GFLOP)" % (k, k, stride, cin, cout, n, h, h, flops / 1e9))
    for name, env in CONFIGS:
        for key in KEYS:
            os.environ.pop(key, None)
        os.environ.update(env)
        plan = N.IgemmPlan()
        xs = x.struct()
        N.check(L.aclgan_plan_conv_fwd(C.byref(layer.desc), C.byref(xs), layer.wptr(0), C.byref(o), C.byref(plan)), "plan")
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        N.check(L.aclgan_igemm_launch_repeat(C.byref(plan), 2, sp), "warm")
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):                      # cold L2 (flushed), single launch
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(L.aclgan_igemm_launch_repeat(C.byref(plan), 1, sp), "launch")
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(L.aclgan_igemm_launch_repeat(C.byref(plan), 10, sp), "launch")     # back to back (operands L2-resident)
        e1.record()
        torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / 10
        print("   %-52s cold %7.1f us %7.1f TF/s | warm %7.1f us %7.1f TF/s" % (
            name, best * 1e3, flops / best / 1e9, warm * 1e3, flops / warm / 1e9))
        if env.get("ACLGAN_SEGK") == "1":
            prof.zero_()
            L.aclgan_igemm_set_prof(prof.data_ptr())
            N.check(L.aclgan_igemm_launch_repeat(C.byref(plan), 1, sp), "launch")
            torch.cuda.synchronize()
            L.aclgan_igemm_set_prof(0)
            pr = prof.view(148, 16).double()
            act = pr[:, 2] > 0
            m = pr[act].mean(0) / 1.9e3      # us at ~1.9 GHz
            print("      role timers (us, mean over issuing CTAs): producer %.1f (waiting %.1f) | MMA thread %.1f (waiting operands "
                  "%.1f, accumulator %.1f) | epilogue %.1f (waiting %.1f; tile fn %.1f = fast %.1f: tmem wait %.1f, chunks %.1f)" % (
                      m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10]))
