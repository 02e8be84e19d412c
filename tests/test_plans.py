"""CPU validation of the host-side plan builders (csrc/plans.cu) by emulating their TMA/MMA/epilogue contract."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

import aclgan_native as N
import emul


def make_act(mem, x_nchw, pad, cs, planes=1, slack=64, mode="reflect"):
    """NCHW float -> padded NHWC bf16 plane(s) with `cs` stored channels (+ zeroed slack behind)."""
    n, c, h, w = x_nchw.shape
    xp = F.pad(x_nchw, (pad,) * 4, mode=mode) if pad > 0 else x_nchw
    buf = torch.zeros(planes, n * (h + 2 * pad) * (w + 2 * pad) * cs + slack, dtype=torch.bfloat16)
    v = torch.zeros(n, h + 2 * pad, w + 2 * pad, cs)
    v[..., :c] = xp.permute(0, 2, 3, 1)
    hi = v.bfloat16()
    buf[0, :v.numel()] = hi.reshape(-1)
    if planes == 2:
        buf[1, :v.numel()] = (v - hi.float()).bfloat16().reshape(-1)
    act = N.Act()
    for p in range(planes):
        act.data[p] = mem.add(buf[p])
    act.planes, act.n, act.h, act.w, act.c, act.pad = planes, n, h, w, cs, pad
    recon = buf[:, :v.numel()].float().sum(0).reshape(n, h + 2 * pad, w + 2 * pad, cs)[..., :c].permute(0, 3, 1, 2)
    return act, buf, recon.double()      # recon = what the tensor cores will effectively see (already padded)


def out_spec(mem, n, h, w, c, kind=N.OUT_F32, pad=0, act=N.ACT_NONE, bias=None, mirror=0):
    hp, wp = h + 2 * pad, w + 2 * pad
    dt = torch.float32 if kind in (N.OUT_F32, N.OUT_F32_ATOMIC) else torch.bfloat16
    planes = 2 if kind == N.OUT_SPLIT else 1
    buf = torch.zeros(planes, n, hp, wp, c, dtype=dt)
    o = N.OutSpec()
    for p in range(planes):
        o.ptr[p] = mem.add(buf[p])
    o.kind, o.act, o.slope, o.mirror = kind, act, 0.2, mirror
    o.off = (pad * wp + pad) * c
    o.sn, o.sy, o.sx, o.sc = hp * wp * c, wp * c, c, 1
    o.N, o.H, o.W, o.C = n, h, w, c
    if bias is not None:
        o.bias = mem.add(bias)
        o.bias_n = bias.numel()
    return o, buf


FWD_CASES = [
    # cin, cout, k, stride, pad, window, n, h, w, planes
    (64, 64, 3, 1, 1, 0, 2, 8, 8, 1),
    (64, 128, 4, 2, 1, 0, 2, 16, 16, 1),
    (128, 64, 5, 1, 2, 0, 1, 8, 16, 1),
    (64, 32, 3, 1, 1, 0, 3, 4, 4, 2),
    (3, 64, 7, 1, 3, 1, 1, 8, 16, 1),
    (6, 16, 4, 2, 1, 1, 2, 8, 8, 1),
    (64, 4, 7, 1, 3, 0, 1, 8, 8, 1),
    (64, 48, 4, 2, 1, 0, 1, 2, 2, 1),
    (64, 64, 5, 1, 2, 0, 1, 3, 128, 1),        # row tiles (128-pixel output rows) in segment mode
    (128, 16, 3, 1, 1, 0, 2, 3, 128, 2),
    # W = 256 (benchmarked geometry): two row tiles per output row; pixel windows / small cout at 256 wide
    (64, 64, 5, 1, 2, 0, 1, 3, 256, 1),
    (3, 64, 7, 1, 3, 1, 1, 4, 256, 1),
    (64, 4, 7, 1, 3, 0, 1, 4, 256, 1),
    (6, 16, 4, 2, 1, 1, 1, 4, 256, 1),
    # vertical-segment window plans (16 x 8 pixel tiles): ragged tile grid, several images, hi/lo planes, 256-wide rows
    (3, 64, 7, 1, 3, 1, 2, 12, 40, 1),
    (3, 64, 7, 1, 3, 1, 1, 8, 32, 2),
    (3, 64, 7, 1, 3, 1, 1, 9, 256, 1),
]


@pytest.mark.parametrize("cin,cout,k,s,pad,window,n,h,w,planes", FWD_CASES)
def test_conv_fwd_plan(cin, cout, k, s, pad, window, n, h, w, planes):
    L = N.lib()
    torch.manual_seed(0)
    mem = emul.Memory()
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) * 0.1
    bias = torch.randn(cout)
    desc = N.ConvDesc(cin, cout, k, s, pad, window)
    cs = (16 if s == 2 else 8) if window else ((cin + 63) // 64) * 64
    act, abuf, xeff = make_act(mem, x, pad, cs, planes)
    wp_hi = emul.pack_weight(desc, wt, False)
    wts = [wp_hi]
    weff = wt.bfloat16().double()
    if planes == 2:
        wts.append(emul.pack_weight(desc, wt - wt.bfloat16().float(), False))
        weff = weff + (wt - wt.bfloat16().float()).bfloat16().double()
    wptr = (C.c_uint64 * 2)(*[mem.add(t) for t in wts] + [0] * (2 - len(wts)))
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    o, obuf = out_spec(mem, n, ho, wo, cout, N.OUT_F32, pad=1, act=N.ACT_LRELU, bias=bias, mirror=1)
    plan = N.IgemmPlan()
    N.check(L.aclgan_plan_conv_fwd(C.byref(desc), C.byref(act), wptr, C.byref(o), C.byref(plan)), "plan fwd")
    assert bool(plan.seg_mode) == (s == 1 and k > 1 and (window != 1 or (wo >= 16 and ho >= 8)))
    # the box-per-tap description (plain kernels) and, when present, the segment description (igemm_seg_kernel)
    for use_seg in ([False, True] if plan.seg_mode else [False]):
        obuf.zero_()
        emul.run_igemm(mem, plan, use_seg=use_seg)
        ref = F.conv2d(xeff, weff, bias.double(), stride=s)
        ref = F.leaky_relu(ref, 0.2)
        ref = F.pad(ref, (1, 1, 1, 1), mode="reflect") if min(ho, wo) > 1 else None
        got = obuf[0].permute(0, 3, 1, 2).double()
        if ref is None:
            ref = F.leaky_relu(F.conv2d(xeff, weff, bias.double(), stride=s), 0.2)
            got = got[:, :, 1:-1, 1:-1]
        tol = 1e-6 if planes == 1 else 2e-4
        assert torch.allclose(got, ref, rtol=tol, atol=tol), (use_seg, float((got - ref).abs().max()))


DGRAD_CASES = [
    # cin, cout, k, stride, pad, n, ho, wo, planes
    (64, 64, 3, 1, 1, 2, 6, 6, 1),
    (64, 128, 4, 2, 1, 1, 4, 4, 1),
    (128, 64, 5, 1, 2, 1, 4, 8, 1),
    (3, 64, 7, 1, 3, 1, 4, 4, 1),
    (64, 64, 4, 2, 1, 2, 2, 2, 2),
    (64, 4, 7, 1, 3, 2, 4, 4, 1),
    (64, 4, 7, 1, 3, 1, 8, 8, 2),
    (64, 64, 5, 1, 2, 1, 2, 256, 1),          # 256-wide output rows (benchmarked geometry)
    (64, 4, 7, 1, 3, 1, 2, 256, 1),
    (64, 4, 7, 1, 3, 2, 9, 21, 1),            # vertical window segments: ragged tile grid, two images
    (64, 3, 7, 1, 3, 1, 10, 18, 2),
]


@pytest.mark.parametrize("cin,cout,k,s,pad,n,ho,wo,planes", DGRAD_CASES)
def test_conv_dgrad_plan(cin, cout, k, s, pad, n, ho, wo, planes):
    L = N.lib()
    torch.manual_seed(1)
    mem = emul.Memory()
    dy = torch.randn(n, cout, ho, wo)
    wt = torch.randn(cout, cin, k, k) * 0.1
    window = 2 if cout <= 8 else (1 if cin <= 8 else 0)   # tiny cout: pixel-window dY (final conv); tiny cin: first conv (fold-mode dgrad)
    desc = N.ConvDesc(cin, cout, k, s, pad, window)
    pz = k - 1 if s == 1 else k // 2 - 1
    cs = 8 if window == 2 else ((cout + 63) // 64) * 64
    act, abuf, dyeff = make_act(mem, dy, pz, cs, planes, mode="constant")
    dyeff = dyeff[:, :, pz:pz + ho, pz:pz + wo] if pz > 0 else dyeff
    wts = [emul.pack_weight(desc, wt, True)]
    weff = wt.bfloat16().double()
    if planes == 2:
        wts.append(emul.pack_weight(desc, wt - wt.bfloat16().float(), True))
        weff = weff + (wt - wt.bfloat16().float()).bfloat16().double()
    wptr = (C.c_uint64 * 2)(*[mem.add(t) for t in wts] + [0] * (2 - len(wts)))
    hp = (ho - 1) * s + k
    wp = (wo - 1) * s + k
    cin_s = ((cin + 15) // 16) * 16
    obuf = torch.zeros(n, hp, wp, cin_s, dtype=torch.float32)
    optr = mem.add(obuf)
    merged = (s == 2 and n % 2 == 0)          # even batch sizes exercise the single-launch (4 phases merged) plan
    # stride-1 layouts: run the segment description of the plan too (pixel-window dY: vertical segments of 16 x 8 tiles)
    seg_pass = (s == 1 and window != 1 and (not window or wo + 2 * (k - 1) >= 16))       # (window 1: fold-mode plan)
    for phase in ([-1] if merged else range(1 if s == 1 else 4)):
        o = N.OutSpec()
        o.ptr[0] = optr
        o.kind, o.act, o.mirror = N.OUT_F32, N.ACT_NONE, 0
        pa, pb = max(phase, 0) >> 1, max(phase, 0) & 1
        o.off = (pa * wp + pb) * cin_s if s == 2 else 0
        o.sn, o.sy, o.sx, o.sc = hp * wp * cin_s, s * wp * cin_s, s * cin_s, 1
        o.N, o.H, o.W, o.C = n, hp // s, wp // s, cin_s
        plan = N.IgemmPlan()
        N.check(L.aclgan_plan_conv_dgrad(C.byref(desc), C.byref(act), wptr, phase, C.byref(o), C.byref(plan)), "plan dgrad")
        assert bool(plan.seg_mode) == seg_pass
        if seg_pass:
            emul.run_igemm(mem, plan, use_seg=False)
            first = obuf.clone()
            obuf.zero_()
        emul.run_igemm(mem, plan)
        if seg_pass:
            assert torch.equal(first, obuf) or torch.allclose(first, obuf, rtol=1e-6, atol=1e-6)
    ref = F.conv_transpose2d(dyeff, weff, stride=s)            # gradient w.r.t. the padded input
    got = obuf.permute(0, 3, 1, 2).double()[:, :cin]
    tol = 1e-6 if planes == 1 else 2e-4
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, rtol=tol, atol=tol), float((got - ref).abs().max())
    assert float(obuf[..., cin:].abs().max()) == 0 if cin_s > cin else True


WGRAD_CASES = [
    # cin, cout, k, stride, pad, window, n, h, w, planes
    (64, 64, 3, 1, 1, 0, 2, 8, 8, 1),
    (64, 128, 4, 2, 1, 0, 2, 16, 16, 1),
    (128, 64, 5, 1, 2, 0, 1, 8, 16, 1),       # swapped roles (cout < 128, cout < cin)
    (256, 128, 3, 1, 1, 0, 1, 4, 4, 2),
    (3, 64, 7, 1, 3, 1, 1, 8, 16, 1),
    (6, 64, 4, 2, 1, 1, 2, 8, 8, 1),
    (64, 4, 7, 1, 3, 2, 1, 8, 8, 1),
    (128, 64, 4, 2, 1, 0, 3, 4, 4, 1),
    (64, 64, 3, 1, 1, 0, 2, 32, 64, 1),       # >= 4096 reduction pixels, 2 chunks: 128-pixel stages
    (3, 64, 7, 1, 3, 1, 1, 64, 64, 1),
    (64, 4, 7, 1, 3, 2, 1, 64, 64, 1),
    # segment mode (stride 1, output rows a multiple of 64 pixels): N-shifted, M-shifted (swapped roles), two planes
    (128, 128, 3, 1, 1, 0, 1, 3, 64, 1),
    (128, 64, 5, 1, 2, 0, 1, 3, 128, 1),
    (64, 128, 3, 1, 1, 0, 2, 2, 64, 2),
    (64, 64, 5, 1, 2, 0, 1, 3, 256, 1),       # W = 256 (benchmarked geometry)
    (64, 4, 7, 1, 3, 2, 1, 4, 256, 1),
    # vertical window segments (16 x 4 pixel blocks): ragged block grid / several images, hi / lo planes, 256-wide rows
    (3, 64, 7, 1, 3, 1, 2, 10, 24, 1),
    (3, 64, 7, 1, 3, 1, 1, 8, 16, 2),
    (3, 64, 7, 1, 3, 1, 1, 5, 256, 1),
    (64, 4, 7, 1, 3, 2, 2, 10, 24, 1),
    (64, 3, 7, 1, 3, 2, 1, 8, 16, 2),
]


@pytest.mark.parametrize("cin,cout,k,s,pad,window,n,h,w,planes", WGRAD_CASES)
def test_conv_wgrad_plan(cin, cout, k, s, pad, window, n, h, w, planes):
    L = N.lib()
    torch.manual_seed(2)
    mem = emul.Memory()
    desc = N.ConvDesc(cin, cout, k, s, pad, window)
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    x = torch.randn(n, cin, h, w)
    dy = torch.randn(n, cout, ho, wo)
    cs_x = (16 if s == 2 else 8) if window == 1 else ((cin + 63) // 64) * 64
    cs_y = 8 if window == 2 else ((cout + 63) // 64) * 64
    pz = (k - 1) if (s == 1) else (k // 2 - 1)
    xact, _, xeff = make_act(mem, x, pad, cs_x, planes)
    yact, _, yeff = make_act(mem, dy, pz, cs_y, planes, mode="constant")
    yeff = yeff[:, :, pz:pz + ho, pz:pz + wo]
    layout = L.aclgan_wgrad_layout(C.byref(desc))
    rows, kt = C.c_int64(), C.c_int64()
    L.aclgan_packed_weight_shape(C.byref(desc), layout, C.byref(rows), C.byref(kt))
    dw = torch.zeros(rows.value * kt.value, dtype=torch.float64)
    plan = N.WgradPlan()
    N.check(L.aclgan_plan_conv_wgrad(C.byref(desc), C.byref(yact), C.byref(xact), mem.add(dw), C.byref(plan)), "plan wgrad")
    seg_on = os.environ.get("ACLGAN_WGRAD_SEG", "1") != "0"
    assert bool(plan.seg_mode) == ((s == 1 and window == 0 and 1 < k <= 7 and wo % 64 == 0 and seg_on) or
                                   (s == 1 and window == 1 and wo >= 16 and ho >= 4) or
                                   (s == 1 and window == 2 and w + 2 * pad >= 16 and ho >= 4))
    emul.run_wgrad(mem, plan)
    got = emul.unpack_wgrad(desc, dw, (cout, cin, k, k))
    ref = torch.nn.grad.conv2d_weight(xeff, (cout, cin, k, k), yeff, stride=s)
    tol = 1e-6 if planes == 1 else 3e-4
    assert torch.allclose(got, ref, rtol=tol, atol=tol * float(ref.abs().max())), float((got - ref).abs().max())


@pytest.mark.parametrize("cout,n,h,w,planes", [(4, 1, 8, 16, 1), (3, 2, 12, 8, 2)])
def test_conv_fwd_fold_plan(monkeypatch, cout, n, h, w, planes):
    """experimental fold mode of the final 7x7 conv (ACLGAN_FOLD=1): filter columns folded into N, diagonal sum in the
    epilogue - packing, plan geometry and epilogue rule validated by CPU emulation"""
    monkeypatch.setenv("ACLGAN_FOLD", "1")
    L = N.lib()
    torch.manual_seed(3)
    mem = emul.Memory()
    cin, k, pad = 64, 7, 3
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) * 0.05
    bias = torch.randn(cout)
    desc = N.ConvDesc(cin, cout, k, 1, pad, N.WINDOW_OUT)
    act, abuf, xeff = make_act(mem, x, pad, 64, planes)
    wts = [emul.pack_weight(desc, wt, False)]
    weff = wt.bfloat16().double()
    if planes == 2:
        wts.append(emul.pack_weight(desc, wt - wt.bfloat16().float(), False))
        weff = weff + (wt - wt.bfloat16().float()).bfloat16().double()
    assert tuple(wts[0].shape) == (64, k * 64)
    wptr = (C.c_uint64 * 2)(*[mem.add(t) for t in wts] + [0] * (2 - len(wts)))
    img = torch.zeros(n, cout, h, w, dtype=torch.float32)
    o = N.OutSpec()
    o.ptr[0] = mem.add(img)
    o.kind, o.act, o.slope, o.mirror, o.off = N.OUT_F32, N.ACT_TANH, 0.2, 0, 0
    o.sn, o.sy, o.sx, o.sc = cout * h * w, w, 1, h * w
    o.N, o.H, o.W, o.C = n, h, w, cout
    o.bias, o.bias_n = mem.add(bias), cout
    plan = N.IgemmPlan()
    N.check(L.aclgan_plan_conv_fwd(C.byref(desc), C.byref(act), wptr, C.byref(o), C.byref(plan)), "plan fwd")
    assert plan.fold == k and plan.tile_step == 120 and plan.num_taps == k and not plan.seg_mode
    emul.run_igemm(mem, plan)
    ref = torch.tanh(F.conv2d(xeff, weff, bias.double()))
    tol = 1e-6 if planes == 1 else 2e-4
    assert torch.allclose(img.double(), ref, rtol=tol, atol=tol), float((img.double() - ref).abs().max())


# Ragged geometries (crop sizes that are not powers of two: the drop-in contract of INTEGRATION.md says "other sizes run with masked
# tiles"): widths that leave partial 128-pixel row tiles, odd heights, partial 16 x 8 window tiles, odd stride-2 inputs.
@pytest.mark.parametrize("cin,cout,k,s,pad,window,n,h,w,planes", [
    (64, 64, 3, 1, 1, 0, 1, 7, 130, 1), (64, 64, 3, 1, 1, 0, 2, 45, 45, 1), (64, 128, 4, 2, 1, 0, 1, 10, 100, 1),
    (128, 64, 5, 1, 2, 0, 1, 6, 72, 1), (3, 64, 7, 1, 3, 1, 1, 13, 67, 1), (64, 4, 7, 1, 3, 0, 1, 9, 100, 1),
    (6, 16, 4, 2, 1, 1, 1, 12, 100, 1)])
def test_conv_fwd_plan_ragged(cin, cout, k, s, pad, window, n, h, w, planes):
    test_conv_fwd_plan(cin, cout, k, s, pad, window, n, h, w, planes)


@pytest.mark.parametrize("cin,cout,k,s,pad,n,ho,wo,planes", [
    (64, 64, 3, 1, 1, 1, 5, 96, 1), (64, 128, 4, 2, 1, 1, 5, 50, 1), (128, 64, 5, 1, 2, 1, 4, 72, 1), (3, 64, 7, 1, 3, 1, 6, 100, 1),
    (64, 64, 3, 1, 1, 2, 9, 45, 2)])
def test_conv_dgrad_plan_ragged(cin, cout, k, s, pad, n, ho, wo, planes):
    test_conv_dgrad_plan(cin, cout, k, s, pad, n, ho, wo, planes)


@pytest.mark.parametrize("cin,cout,k,s,pad,window,n,h,w,planes", [
    (64, 64, 3, 1, 1, 0, 1, 5, 96, 1), (64, 128, 4, 2, 1, 0, 1, 10, 100, 1), (3, 64, 7, 1, 3, 1, 1, 10, 100, 1),
    (64, 4, 7, 1, 3, 0, 1, 9, 100, 1), (64, 64, 3, 1, 1, 0, 2, 45, 45, 1)])
def test_conv_wgrad_plan_ragged(cin, cout, k, s, pad, window, n, h, w, planes):
    test_conv_wgrad_plan(cin, cout, k, s, pad, window, n, h, w, planes)


@pytest.mark.parametrize("cin,cout,n,w,planes", [(128, 64, 2, 64, 1), (64, 64, 4, 128, 2)])
def test_conv_wgrad_strip_plan_is_box_per_tap_at_64_pixel_multiples(monkeypatch, cin, cout, n, w, planes):
    """the column strips of the sub-pixel up blocks (2 rows x 2H - 4 pixels, 5x5) need a box-per-tap weight-gradient plan (their
    taps are stored transposed, engine.conv_wgrad(transpose_taps=True)); for H = 34, 66, ... the strip is a multiple of 64 pixels
    long, where the builder prefers a segment plan: the engine switches that choice off for the call.  Same switch, same geometry
    here, result against torch."""
    import engine as E
    with E._plan_env("ACLGAN_WGRAD_SEG", "0"):
        assert os.environ["ACLGAN_WGRAD_SEG"] == "0"
        test_conv_wgrad_plan(cin, cout, 5, 1, 2, 0, n, 3, w, planes)      # (3 rows: the test helper reflect-pads by 2)
    assert "ACLGAN_WGRAD_SEG" not in os.environ


@pytest.mark.parametrize("w", [60, 64])
def test_conv_wgrad_transposed_taps(w):
    """engine.conv_wgrad(transpose_taps=True), step for step on the CPU: box-per-tap plan (segment choice off), tap slot (kh, kw)
    redirected to (kw, kh), emulated launch - the stored gradient is the weight gradient with its two filter axes swapped.
    w = 64 is the strip length at which the builder would otherwise choose a segment plan."""
    import engine as E
    L = N.lib()
    torch.manual_seed(5)
    cin, cout, k, n, h = 128, 64, 5, 2, 3
    mem = emul.Memory()
    desc = N.ConvDesc(cin, cout, k, 1, 2, 0)
    x = torch.randn(n, cin, h, w)
    dy = torch.randn(n, cout, h, w)
    xact, _, xeff = make_act(mem, x, 2, cin, 1)
    yact, _, yeff = make_act(mem, dy, k - 1, cout, 1, mode="constant")
    yeff = yeff[:, :, k - 1:k - 1 + h, k - 1:k - 1 + w]
    layout = L.aclgan_wgrad_layout(C.byref(desc))
    rows, kt = C.c_int64(), C.c_int64()
    L.aclgan_packed_weight_shape(C.byref(desc), layout, C.byref(rows), C.byref(kt))
    dw = torch.zeros(rows.value * kt.value, dtype=torch.float64)
    plan = N.WgradPlan()
    with E._plan_env("ACLGAN_WGRAD_SEG", "0"):
        N.check(L.aclgan_plan_conv_wgrad(C.byref(desc), C.byref(yact), C.byref(xact), mem.add(dw), C.byref(plan)), "plan wgrad")
    assert plan.seg_mode == 0 and plan.num_taps == k * k
    for t in range(k * k):
        plan.tap_out[t] = (t % k) * k + t // k
    emul.run_wgrad(mem, plan)
    got = emul.unpack_wgrad(desc, dw, (cout, cin, k, k))
    ref = torch.nn.grad.conv2d_weight(xeff, (cout, cin, k, k), yeff, stride=1).transpose(2, 3)
    assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6 * float(ref.abs().max())), float((got - ref).abs().max())
