"""-m gpu: the tcgen05/TMA implicit-GEMM kernels through the C ABI vs torch convolutions in fp64."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import aclgan_native as N
import gpu_util as G

pytestmark = pytest.mark.gpu

FWD_CASES = [
    # cin, cout, k, stride, pad, window, n, h, w, planes
    (64, 64, 3, 1, 1, 0, 2, 8, 8, 1),
    (256, 256, 3, 1, 1, 0, 2, 64, 64, 1),
    (64, 128, 4, 2, 1, 0, 2, 32, 32, 1),
    (256, 128, 5, 1, 2, 0, 1, 32, 32, 1),
    (256, 512, 4, 2, 1, 0, 2, 16, 16, 1),
    (64, 32, 3, 1, 1, 0, 3, 4, 4, 2),
    (256, 256, 3, 1, 1, 0, 1, 16, 16, 2),
    (3, 64, 7, 1, 3, 1, 1, 32, 64, 1),
    (6, 64, 4, 2, 1, 1, 2, 32, 32, 1),
    (64, 4, 7, 1, 3, 0, 1, 32, 32, 1),
    (64, 48, 4, 2, 1, 0, 1, 2, 2, 1),
    # W = 256 (the benchmarked geometry): two 128-pixel tiles per output row, 258/260/262-wide padded planes
    (128, 64, 5, 1, 2, 0, 1, 8, 256, 1),       # 5x5 128->64: one-CTA N = 64 segment path
    (128, 64, 5, 1, 2, 0, 1, 4, 256, 2),
    (64, 4, 7, 1, 3, 0, 1, 8, 256, 1),         # final 7x7 64->4
    (64, 4, 7, 1, 3, 0, 1, 4, 256, 2),
    (3, 64, 7, 1, 3, 1, 1, 8, 256, 1),         # first 7x7 3->64: pixel windows at W = 256
    (6, 64, 4, 2, 1, 1, 1, 16, 256, 1),        # discriminator first conv on a 256-wide pair
    (256, 256, 3, 1, 1, 0, 1, 4, 256, 1),
    # vertical window segments + resident weights: ragged 16 x 8 tile grid over several images, hi / lo planes
    (3, 64, 7, 1, 3, 1, 3, 20, 40, 1),
    (3, 64, 7, 1, 3, 1, 2, 16, 32, 2),
]


@pytest.mark.parametrize("msub", [1, 2, "pair"])
@pytest.mark.parametrize("cin,cout,k,s,pad,window,n,h,w,planes", FWD_CASES)
def test_conv_fwd(monkeypatch, msub, cin, cout, k, s, pad, window, n, h, w, planes):
    # kernel variants: 128- / 256-pixel work items of the 1-CTA kernel, or the CTA-pair (cta_group::2) kernel
    monkeypatch.setenv("ACLGAN_IGEMM_PAIR", "1" if msub == "pair" else "0")
    monkeypatch.setenv("ACLGAN_IGEMM_MSUB", "1" if msub == "pair" else str(msub))
    L = N.lib()
    torch.manual_seed(0)
    x = torch.randn(n, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)
    bias = torch.randn(cout, device="cuda")
    desc = N.ConvDesc(cin, cout, k, s, pad, window)
    cs = (16 if s == 2 else 8) if window else ((cin + 63) // 64) * 64
    act, abuf, xeff = G.make_act(x, pad, cs, planes)
    packed, weff = G.pack_weights(desc, wt, False, planes)
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    o, obuf = G.out_spec(n, ho, wo, cout, N.OUT_F32, pad=1, act=N.ACT_LRELU, bias=bias, mirror=1)
    plan = N.IgemmPlan()
    N.check(L.aclgan_plan_conv_fwd(C.byref(desc), C.byref(act), G.wptr(packed), C.byref(o), C.byref(plan)), "plan")
    N.check(L.aclgan_igemm_launch(C.byref(plan), G.stream_ptr()), "launch")
    torch.cuda.synchronize()
    ref = F.leaky_relu(F.conv2d(xeff, weff.cuda(), bias.double(), stride=s), 0.2)
    got = obuf[0].permute(0, 3, 1, 2).double()
    if min(ho, wo) > 1:
        ref = F.pad(ref, (1, 1, 1, 1), mode="reflect")
    else:
        got = got[:, :, 1:-1, 1:-1]
    err = G.rel_err(got, ref)
    # operands are exactly representable; only the fp32 accumulation order differs (x3 also drops lo*lo)
    assert err < (2e-5 if planes == 1 else 5e-5), err  # fp32 accumulation over K up to 6400


def test_conv_fwd_bf16_store_and_split():
    L = N.lib()
    torch.manual_seed(1)
    n, cin, cout, h = 2, 128, 128, 16
    x = torch.randn(n, cin, h, h, device="cuda")
    wt = torch.randn(cout, cin, 3, 3) * 0.03
    desc = N.ConvDesc(cin, cout, 3, 1, 1, 0)
    act, _, xeff = G.make_act(x, 1, cin, 2)
    packed, weff = G.pack_weights(desc, wt, False, 2)
    ref = F.conv2d(xeff, weff.cuda())
    for kind in (N.OUT_BF16, N.OUT_SPLIT):
        o, obuf = G.out_spec(n, h, h, cout, kind, pad=2, mirror=2)
        plan = N.IgemmPlan()
        N.check(L.aclgan_plan_conv_fwd(C.byref(desc), C.byref(act), G.wptr(packed), C.byref(o), C.byref(plan)), "plan")
        N.check(L.aclgan_igemm_launch(C.byref(plan), G.stream_ptr()), "launch")
        torch.cuda.synchronize()
        got = obuf.float().sum(0).permute(0, 3, 1, 2).double()
        err = G.rel_err(got, F.pad(ref, (2, 2, 2, 2), mode="reflect"))
        assert err < (4e-3 if kind == N.OUT_BF16 else 3e-5), (kind, err)


DGRAD_CASES = [
    (64, 64, 3, 1, 1, 2, 6, 6, 1),
    (256, 256, 3, 1, 1, 2, 64, 64, 1),
    (64, 128, 4, 2, 1, 1, 16, 16, 1),
    (256, 128, 5, 1, 2, 1, 32, 32, 1),
    (3, 64, 7, 1, 3, 1, 16, 16, 1),
    (64, 64, 4, 2, 1, 2, 2, 2, 2),
    (64, 4, 7, 1, 3, 2, 8, 8, 1),
    (64, 4, 7, 1, 3, 2, 8, 8, 2),
    (64, 3, 7, 1, 3, 1, 32, 32, 1),
    (256, 128, 5, 1, 2, 2, 128, 128, 1),      # cin 256 -> N = 256 pairs, 5 taps per stage
    (128, 64, 5, 1, 2, 2, 128, 128, 1),       # cin 128 -> N = 128 pairs
    (128, 64, 5, 1, 2, 2, 128, 128, 2),
    (64, 64, 3, 1, 1, 2, 128, 128, 2),
    # 256-wide output rows (benchmarked geometry)
    (128, 64, 5, 1, 2, 1, 8, 256, 1),
    (64, 4, 7, 1, 3, 1, 8, 256, 1),
    (64, 4, 7, 1, 3, 1, 4, 256, 2),
    (3, 64, 7, 1, 3, 1, 8, 256, 1),
    (64, 128, 4, 2, 1, 1, 8, 128, 1),          # stride-2 data gradient of a 256-wide input
]


@pytest.mark.parametrize("msub", [1, 2, "pair"])
@pytest.mark.parametrize("cin,cout,k,s,pad,n,ho,wo,planes", DGRAD_CASES)
def test_conv_dgrad(monkeypatch, msub, cin, cout, k, s, pad, n, ho, wo, planes):
    monkeypatch.setenv("ACLGAN_IGEMM_PAIR", "1" if msub == "pair" else "0")
    monkeypatch.setenv("ACLGAN_IGEMM_MSUB", "1" if msub == "pair" else str(msub))
    L = N.lib()
    torch.manual_seed(1)
    dy = torch.randn(n, cout, ho, wo, device="cuda")
    wt = torch.randn(cout, cin, k, k) * 0.05
    window = 2 if cout <= 8 else (1 if cin <= 8 else 0)   # tiny cout: pixel-window dY (final conv); tiny cin: first conv (fold-mode dgrad)
    desc = N.ConvDesc(cin, cout, k, s, pad, window)
    pz = k - 1 if s == 1 else k // 2 - 1
    cs = 8 if window == 2 else ((cout + 63) // 64) * 64
    act, abuf, dyeff = G.make_act(dy, pz, cs, planes, mode="constant")
    if pz > 0:
        dyeff = dyeff[:, :, pz:pz + ho, pz:pz + wo]
    packed, weff = G.pack_weights(desc, wt, True, planes)
    hp, wp = (ho - 1) * s + k, (wo - 1) * s + k
    cin_s = ((cin + 15) // 16) * 16
    obuf = torch.zeros(n, hp, wp, cin_s, dtype=torch.float32, device="cuda")
    merged = (s == 2 and n % 2 == 0)          # even batch sizes exercise the single-launch (4 phases merged) plan
    for phase in ([-1] if merged else range(1 if s == 1 else 4)):
        o = N.OutSpec()
        o.ptr[0] = obuf.data_ptr()
        o.kind, o.act, o.mirror = N.OUT_F32, N.ACT_NONE, 0
        pa, pb = max(phase, 0) >> 1, max(phase, 0) & 1
        o.off = (pa * wp + pb) * cin_s if s == 2 else 0
        o.sn, o.sy, o.sx, o.sc = hp * wp * cin_s, s * wp * cin_s, s * cin_s, 1
        o.N, o.H, o.W, o.C = n, hp // s, wp // s, cin_s
        plan = N.IgemmPlan()
        N.check(L.aclgan_plan_conv_dgrad(C.byref(desc), C.byref(act), G.wptr(packed), phase, C.byref(o), C.byref(plan)), "plan")
        N.check(L.aclgan_igemm_launch(C.byref(plan), G.stream_ptr()), "launch")
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(dyeff, weff.cuda(), stride=s)
    got = obuf.permute(0, 3, 1, 2).double()[:, :cin]
    err = G.rel_err(got, ref)
    assert err < (2e-5 if planes == 1 else 5e-5), err  # fp32 accumulation over K up to 6400


@pytest.mark.parametrize("msub", [1, 2, "pair"])
def test_conv_fwd_throughput_report(capsys, monkeypatch, msub):
    """not an assertion on speed - prints the achieved TFLOP/s of the dominant 3x3 256->256 layer"""
    monkeypatch.setenv("ACLGAN_IGEMM_PAIR", "1" if msub == "pair" else "0")
    monkeypatch.setenv("ACLGAN_IGEMM_MSUB", "1" if msub == "pair" else str(msub))
    L = N.lib()
    torch.manual_seed(0)
    n, c, h = 8, 256, 64
    x = torch.randn(n, c, h, h, device="cuda")
    wt = torch.randn(c, c, 3, 3) * 0.02
    desc = N.ConvDesc(c, c, 3, 1, 1, 0)
    act, _, _ = G.make_act(x, 1, c, 1)
    packed, _ = G.pack_weights(desc, wt, False, 1)
    o, obuf = G.out_spec(n, h, h, c, N.OUT_BF16, pad=1, mirror=1)
    plan = N.IgemmPlan()
    N.check(L.aclgan_plan_conv_fwd(C.byref(desc), C.byref(act), G.wptr(packed), C.byref(o), C.byref(plan)), "plan")
    N.check(L.aclgan_igemm_launch_repeat(C.byref(plan), 3, G.stream_ptr()), "launch")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 20
    N.check(L.aclgan_igemm_launch_repeat(C.byref(plan), iters, G.stream_ptr()), "launch")   # device-side timing
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * n * h * h * c * c * 9
    with capsys.disabled():
        print("\n[igemm 3x3 256->256 bs8 64x64 msub=%s] %.3f ms  %.1f TFLOP/s" % (msub, ms, flops / ms / 1e9))


WGRAD_CASES = [
    # cin, cout, k, stride, pad, window, n, h, w, planes
    (64, 64, 3, 1, 1, 0, 2, 8, 8, 1),
    (256, 256, 3, 1, 1, 0, 4, 64, 64, 1),
    (64, 128, 4, 2, 1, 0, 2, 32, 32, 1),
    (128, 64, 5, 1, 2, 0, 1, 32, 32, 1),
    (256, 512, 4, 2, 1, 0, 2, 16, 16, 1),
    (256, 128, 3, 1, 1, 0, 1, 8, 8, 2),
    (3, 64, 7, 1, 3, 1, 1, 32, 32, 1),
    (6, 64, 4, 2, 1, 1, 2, 32, 32, 1),
    (64, 4, 7, 1, 3, 2, 1, 32, 32, 1),
    (128, 64, 4, 2, 1, 0, 3, 4, 4, 1),
    (64, 4, 7, 1, 3, 2, 2, 8, 8, 1),
    (64, 4, 7, 1, 3, 2, 2, 8, 8, 2),
    (64, 128, 4, 2, 1, 0, 1, 32, 32, 2),
    (16, 32, 4, 2, 1, 0, 2, 32, 32, 2),
    (64, 128, 4, 2, 1, 0, 1, 16, 16, 1),
    # segment mode (stride 1, 64-pixel output rows)
    (128, 128, 3, 1, 1, 0, 2, 16, 64, 1),
    (256, 128, 5, 1, 2, 0, 1, 8, 128, 1),
    (128, 64, 5, 1, 2, 0, 1, 8, 64, 1),
    (64, 128, 3, 1, 1, 0, 2, 4, 64, 2),
    # >= 4096 reduction pixels and narrow operands: 128-pixel pipeline stages
    (64, 64, 3, 1, 1, 0, 2, 64, 64, 1),
    (64, 64, 3, 1, 1, 0, 2, 64, 64, 2),
    (128, 64, 5, 1, 2, 0, 2, 64, 64, 1),
    (3, 64, 7, 1, 3, 1, 2, 64, 64, 1),
    (6, 64, 4, 2, 1, 1, 2, 128, 128, 1),
    (64, 4, 7, 1, 3, 2, 2, 64, 64, 1),
    # W = 256 (benchmarked geometry)
    (128, 64, 5, 1, 2, 0, 1, 8, 256, 1),
    (64, 4, 7, 1, 3, 2, 1, 8, 256, 1),
    (64, 4, 7, 1, 3, 2, 1, 4, 256, 2),
    (3, 64, 7, 1, 3, 1, 1, 8, 256, 1),
    (6, 64, 4, 2, 1, 1, 1, 16, 256, 1),
    (64, 128, 4, 2, 1, 0, 1, 16, 256, 1),
    # vertical window segments (merged taps): ragged 16 x 4 block grid, hi / lo planes, both window sides
    (3, 64, 7, 1, 3, 1, 3, 10, 40, 1),
    (3, 64, 7, 1, 3, 1, 2, 16, 32, 2),
    (64, 4, 7, 1, 3, 2, 3, 10, 40, 1),
    (64, 3, 7, 1, 3, 2, 2, 16, 32, 2),
]


@pytest.mark.parametrize("cin,cout,k,s,pad,window,n,h,w,planes", WGRAD_CASES)
def test_conv_wgrad(cin, cout, k, s, pad, window, n, h, w, planes):
    import emul
    L = N.lib()
    torch.manual_seed(2)
    desc = N.ConvDesc(cin, cout, k, s, pad, window)
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    x = torch.randn(n, cin, h, w, device="cuda")
    dy = torch.randn(n, cout, ho, wo, device="cuda")
    cs_x = (16 if s == 2 else 8) if window == 1 else ((cin + 63) // 64) * 64
    cs_y = 8 if window == 2 else ((cout + 63) // 64) * 64
    pz = (k - 1) if (s == 1) else (k // 2 - 1)
    xact, xbuf, xeff = G.make_act(x, pad, cs_x, planes)
    yact, ybuf, yeff = G.make_act(dy, pz, cs_y, planes, mode="constant")
    yeff = yeff[:, :, pz:pz + ho, pz:pz + wo]
    layout = L.aclgan_wgrad_layout(C.byref(desc))
    rows, kt = C.c_int64(), C.c_int64()
    L.aclgan_packed_weight_shape(C.byref(desc), layout, C.byref(rows), C.byref(kt))
    dw = torch.zeros(rows.value * kt.value, dtype=torch.float32, device="cuda")
    plan = N.WgradPlan()
    N.check(L.aclgan_plan_conv_wgrad(C.byref(desc), C.byref(yact), C.byref(xact), dw.data_ptr(), C.byref(plan)), "plan")
    N.check(L.aclgan_wgrad_launch(C.byref(plan), G.stream_ptr()), "launch")
    torch.cuda.synchronize()
    got = emul.unpack_wgrad(desc, dw, (cout, cin, k, k)).double()
    ref = torch.nn.grad.conv2d_weight(xeff, (cout, cin, k, k), yeff, stride=s)
    err = G.rel_err(got, ref)
    assert err < (5e-6 if planes == 1 else 5e-5), err


def test_conv_wgrad_throughput_report(capsys):
    L = N.lib()
    torch.manual_seed(0)
    n, c, h = 8, 256, 64
    desc = N.ConvDesc(c, c, 3, 1, 1, 0)
    xact, xbuf, _ = G.make_act(torch.randn(n, c, h, h, device="cuda"), 1, c, 1)
    yact, ybuf, _ = G.make_act(torch.randn(n, c, h, h, device="cuda"), 2, c, 1, mode="constant")
    dw = torch.zeros(c * 9 * c, dtype=torch.float32, device="cuda")
    plan = N.WgradPlan()
    N.check(L.aclgan_plan_conv_wgrad(C.byref(desc), C.byref(yact), C.byref(xact), dw.data_ptr(), C.byref(plan)), "plan")
    for _ in range(3):
        N.check(L.aclgan_wgrad_launch(C.byref(plan), G.stream_ptr()), "launch")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        N.check(L.aclgan_wgrad_launch(C.byref(plan), G.stream_ptr()), "launch")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    with capsys.disabled():
        print("\n[wgrad 3x3 256->256 bs8 64x64] %.3f ms  %.1f TFLOP/s  (ksplit %d)" % (
            ms, 2.0 * n * h * h * c * c * 9 / ms / 1e9, plan.ksplit))
