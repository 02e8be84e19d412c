"""-m gpu: fused conv blocks (conv + norm + act + residual + upsample + reflect pad), forward and backward,
against a plain torch fp64 restatement on the same operands (tolerances: fp32x3 parity mode 2e-4 / bf16 mode 3e-2)."""
import pytest
import torch
import torch.nn.functional as F

import aclgan_native as N
import engine as E

pytestmark = pytest.mark.gpu


def plane_from_nchw(eng, x, pad):
    n, c, h, w = x.shape
    a = E.ActT(eng, n, h, w, c, pad)
    v = F.pad(x.float(), (pad,) * 4, mode="reflect").permute(0, 2, 3, 1) if pad else x.float().permute(0, 2, 3, 1)
    full = torch.zeros((n, h + 2 * pad, w + 2 * pad, a.c), device=x.device)
    full[..., :c] = v
    hi = full.bfloat16()
    a.buf[0, :a.numel] = hi.reshape(-1)
    if a.planes == 2:
        a.buf[1, :a.numel] = (full - hi.float()).bfloat16().reshape(-1)
    return a


def ref_block(x, w, b, stride, pad, norm, act, res, upsample, out_pad, adain, ln):
    y = F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), w, b, stride=stride)
    if norm == N.NORM_IN or norm == N.NORM_ADAIN:
        mu = y.mean((2, 3), keepdim=True)
        var = y.var((2, 3), unbiased=False, keepdim=True)
        y = (y - mu) / torch.sqrt(var + 1e-5)
        if norm == N.NORM_ADAIN:
            y = y * adain[0][:, :, None, None] + adain[1][:, :, None, None]
    elif norm == N.NORM_LN:
        flat = y.reshape(y.shape[0], -1)
        mu = flat.mean(1).view(-1, 1, 1, 1)
        sd = flat.std(1).view(-1, 1, 1, 1)
        y = (y - mu) / (sd + 1e-5) * ln[0].view(1, -1, 1, 1) + ln[1].view(1, -1, 1, 1)
    if act == N.ACT_RELU:
        y = torch.relu(y)
    elif act == N.ACT_LRELU:
        y = F.leaky_relu(y, 0.2)
    if res is not None:
        y = y + res
    if upsample == 2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    return F.pad(y, (out_pad,) * 4, mode="reflect") if out_pad else y


CASES = [
    # cin, cout, k, stride, pad, norm, act, res, upsample, out_pad, n, h
    (64, 64, 3, 1, 1, N.NORM_IN, N.ACT_RELU, False, 1, 1, 2, 16),
    (64, 64, 3, 1, 1, N.NORM_ADAIN, N.ACT_NONE, True, 2, 2, 2, 8),
    (64, 128, 4, 2, 1, N.NORM_IN, N.ACT_RELU, False, 1, 1, 2, 16),
    (128, 64, 5, 1, 2, N.NORM_LN, N.ACT_RELU, False, 2, 2, 2, 8),
    (128, 64, 5, 1, 2, N.NORM_LN, N.ACT_RELU, False, 1, 3, 1, 8),
    (64, 128, 4, 2, 1, N.NORM_NONE, N.ACT_LRELU, False, 1, 1, 2, 16),
    (64, 64, 4, 2, 1, N.NORM_NONE, N.ACT_RELU, False, 1, 0, 2, 8),
    (32, 16, 3, 1, 1, N.NORM_ADAIN, N.ACT_RELU, False, 1, 1, 2, 8),      # channel counts below one 64-chunk
    # larger planes: several CTAs per image in the reductions, several tiles per image row
    (64, 64, 3, 1, 1, N.NORM_IN, N.ACT_RELU, False, 1, 1, 2, 64),
    (64, 128, 4, 2, 1, N.NORM_IN, N.ACT_RELU, False, 1, 1, 2, 64),
    (64, 64, 4, 2, 1, N.NORM_NONE, N.ACT_LRELU, False, 1, 1, 1, 64),
    (64, 64, 5, 1, 2, N.NORM_LN, N.ACT_RELU, False, 2, 2, 1, 32),
    # 128-pixel output rows: row tiles of the segment kernels, statistics fused in the epilogue, segment weight gradient
    (128, 64, 5, 1, 2, N.NORM_LN, N.ACT_RELU, False, 1, 3, 2, 128),
    (64, 128, 3, 1, 1, N.NORM_ADAIN, N.ACT_RELU, False, 2, 2, 1, 128),
    # the benchmarked 256x256 decoder tail: 5x5 256->128 (output upsampled to 256, pad 2) and 5x5 128->64 on the 256^2 plane
    # (out_pad 3 = the final 7x7's reflect width), 4x4 s2 64->128 from 256 wide
    (256, 128, 5, 1, 2, N.NORM_LN, N.ACT_RELU, False, 2, 2, 1, 128),
    (128, 64, 5, 1, 2, N.NORM_LN, N.ACT_RELU, False, 1, 3, 1, 256),
    (64, 128, 4, 2, 1, N.NORM_IN, N.ACT_RELU, False, 1, 1, 1, 256),
]


@pytest.mark.parametrize("precision", ["fp32x3", "bf16"])
@pytest.mark.parametrize("cin,cout,k,stride,pad,norm,act,use_res,upsample,out_pad,n,h", CASES)
def test_conv_block(precision, cin, cout, k, stride, pad, norm, act, use_res, upsample, out_pad, n, h):
    _block_case(precision, cin, cout, k, stride, pad, norm, act, use_res, upsample, out_pad, n, h, fuse=False)


FUSED_CASES = [c for c in CASES if c[5] != N.NORM_NONE and c[11] <= 128][:9]


@pytest.mark.parametrize("precision", ["fp32x3", "bf16"])
@pytest.mark.parametrize("cin,cout,k,stride,pad,norm,act,use_res,upsample,out_pad,n,h", FUSED_CASES)
def test_conv_block_fused_finalize(precision, cin, cout, k, stride, pad, norm, act, use_res, upsample, out_pad, n, h):
    """the opt-in one-launch forms (aclgan_norm_finalize_apply / aclgan_norm_bwd_finalize_apply, ACLGAN_FUSE_FINALIZE=1): every
    CTA of the row kernels derives the coefficients of its own channels - same arithmetic, same results"""
    _block_case(precision, cin, cout, k, stride, pad, norm, act, use_res, upsample, out_pad, n, h, fuse=True)


def _block_case(precision, cin, cout, k, stride, pad, norm, act, use_res, upsample, out_pad, n, h, fuse):
    eng = E.Engine(precision)
    eng.fuse_finalize = fuse
    tol = 2e-4 if precision == "fp32x3" else 4e-2
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(n, cin, h, h, device=dev)
    w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5)
    b = torch.nn.Parameter(torch.randn(cout, device=dev) * 0.1)
    ho = (h + 2 * pad - k) // stride + 1
    res = torch.randn(n, cout, ho, ho, device=dev) if use_res else None
    adain_w = torch.rand(n, cout, device=dev) + 0.5
    adain_b = torch.randn(n, cout, device=dev)
    gamma, beta = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)

    arena = E.GradArena(eng.device)
    layer = E.ConvLayer(eng, arena, w, b, stride, pad)
    ln_off = (arena.reserve(cout), arena.reserve(cout))
    arena.finalize()
    xa = plane_from_nchw(eng, x, pad)
    xa.requires_grad = True
    ra = plane_from_nchw(eng, res, 1) if use_res else None
    if ra is not None:
        ra.requires_grad = True
    # AdaIN parameters / gradients as strided column slices of one row per sample (how AdaINGen.decode passes them)
    ap = torch.cat((adain_b, adain_w), 1)
    d_ap = torch.zeros_like(ap)
    got_adain = dict(dw=d_ap[:, cout:], db=d_ap[:, :cout])
    adain = (ap[:, cout:], ap[:, :cout], got_adain["dw"], got_adain["db"]) if norm == N.NORM_ADAIN else None
    ln = (gamma, beta, arena.view(ln_off[0], cout), arena.view(ln_off[1], cout)) if norm == N.NORM_LN else None
    tape = E.Tape()
    out = eng.conv_block(tape, layer, xa, norm=norm, act=act, out_pad=out_pad, upsample=upsample, res=ra,
                         adain=adain, ln=ln)
    torch.cuda.synchronize()

    # ---- reference in fp64 on the operands the kernels actually see (bf16 / hi+lo rounded)
    def eff(t, planes):
        hi = t.detach().float().bfloat16()
        return (hi.double() + ((t.detach().float() - hi.float()).bfloat16().double() if planes == 2 else 0))
    P = eng.prec.planes
    x64 = eff(x, P).requires_grad_(True)
    w64 = eff(w, P).requires_grad_(True)
    b64 = b.detach().double().requires_grad_(True)
    r64 = eff(res, P).requires_grad_(True) if use_res else None
    aw, ab = adain_w.double().requires_grad_(True), adain_b.double().requires_grad_(True)
    g64, be64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = ref_block(x64, w64, b64, stride, pad, norm, act, r64, upsample, out_pad, (aw, ab), (g64, be64))

    p = out.pad
    full = out.buf[:, :out.numel].float().sum(0).view(n, out.h + 2 * p, out.w + 2 * p, out.c)
    got = full[..., :cout].permute(0, 3, 1, 2).double()
    err = float((got - ref).norm() / ref.norm())
    assert err < tol, ("forward", err)
    if out.c > cout:
        assert float(full[..., cout:].abs().max()) == 0.0

    # ---- backward: random gradient on the padded output plane
    G = torch.randn_like(ref)
    gp = torch.zeros((n, out.h + 2 * p, out.w + 2 * p, out.c), device=dev, dtype=eng.prec.dtype)
    gp[..., :cout] = G.permute(0, 2, 3, 1).to(eng.prec.dtype)
    Geff = gp[..., :cout].double().permute(0, 3, 1, 2)
    out.gp = gp
    tape.backward()
    torch.cuda.synchronize()
    (ref * Geff).sum().backward()

    def rel(a, bb):
        return float((a.double() - bb).norm() / (bb.norm() + 1e-30))
    # gradient w.r.t. the input: fold my padded-plane gradient with the adjoint of reflect padding
    gx = xa.gp[..., :cin].double().permute(0, 3, 1, 2).requires_grad_(False)
    probe = torch.zeros_like(x64).requires_grad_(True)
    (F.pad(probe, (pad,) * 4, mode="reflect") * gx).sum().backward()
    btol = tol * 3
    if h >= 128 and act != N.ACT_NONE and precision == "fp32x3":
        # millions of units: a handful of ReLU pre-activations lie within the fp32x3 rounding distance (1e-5) of zero and
        # flip against the fp64 reference; each flip moves these relative L2 errors by ~1e-4 (tests/test_gpu_step.py docstring)
        btol = 3e-3
    assert rel(probe.grad, x64.grad) < btol, ("dgrad", rel(probe.grad, x64.grad))
    gw, gb = layer.grad_views()
    assert rel(gw, w64.grad) < btol, ("wgrad", rel(gw, w64.grad))
    if norm == N.NORM_NONE:
        assert rel(gb, b64.grad) < btol, ("bgrad", rel(gb, b64.grad))
    if use_res:
        assert rel(ra.gr[..., :cout].permute(0, 3, 1, 2), r64.grad) < btol, "res grad"
    if norm == N.NORM_ADAIN:
        e_w, e_b = rel(got_adain["dw"], aw.grad), rel(got_adain["db"], ab.grad)
        assert e_w < btol and e_b < btol, ("adain grads", e_w, e_b)
    if norm == N.NORM_LN:
        e_g, e_b = rel(ln[2], g64.grad), rel(ln[3], be64.grad)
        assert e_g < btol and e_b < btol, ("ln grads", e_g, e_b)


@pytest.mark.parametrize("precision", ["fp32x3", "bf16"])
def test_image_io_and_final_conv(precision):
    """image pack (+concat) -> window conv ; 64 -> 4 tanh conv to an NCHW image ; both with backward"""
    eng = E.Engine(precision)
    tol = 2e-4 if precision == "fp32x3" else 4e-2
    torch.manual_seed(1)
    dev = "cuda"
    n, h = 2, 16
    P = eng.prec.planes

    def eff(t):
        hi = t.detach().float().bfloat16()
        return hi.double() + ((t.detach().float() - hi.float()).bfloat16().double() if P == 2 else 0)

    # ---- first conv of a discriminator on a concatenated pair (6 channels, 4x4 stride 2)
    a_img = torch.rand(n, 3, h, h, device=dev) * 2 - 1
    b_img = torch.rand(n, 3, h, h, device=dev) * 2 - 1
    w = torch.nn.Parameter(torch.randn(64, 6, 4, 4, device=dev) * 0.1)
    b = torch.nn.Parameter(torch.randn(64, device=dev) * 0.1)
    arena = E.GradArena(eng.device)
    layer = E.ConvLayer(eng, arena, w, b, 2, 1, N.WINDOW_IN)
    w2 = torch.nn.Parameter(torch.randn(4, 64, 7, 7, device=dev) * 0.02)
    b2 = torch.nn.Parameter(torch.randn(4, device=dev) * 0.1)
    layer2 = E.ConvLayer(eng, arena, w2, b2, 1, 3, N.WINDOW_OUT)
    arena.finalize()
    tape = E.Tape()
    ia, ib = E.ImgT(a_img), E.ImgT(b_img, requires_grad=True)
    xp = eng.pack_image(tape, ia, 1, 16, ib)
    out = eng.conv_block(tape, layer, xp, norm=N.NORM_NONE, act=N.ACT_LRELU, out_pad=3)
    img = eng.conv_to_image(tape, layer2, out)
    torch.cuda.synchronize()

    a64, b64i = eff(a_img), eff(b_img).requires_grad_(True)
    w64, bb64 = eff(w).requires_grad_(True), b.detach().double().requires_grad_(True)
    w264, b264 = eff(w2).requires_grad_(True), b2.detach().double().requires_grad_(True)
    y = F.leaky_relu(F.conv2d(F.pad(torch.cat((a64, b64i), 1), (1,) * 4, mode="reflect"), w64, bb64, stride=2), 0.2)
    y_seen = y
    if P == 1:
        y_seen = y + (y.detach().float().bfloat16().double() - y.detach())      # the next conv reads the bf16 plane
    ref = torch.tanh(F.conv2d(F.pad(y_seen, (3,) * 4, mode="reflect"), w264, b264))
    err = float((img.t.double() - ref).norm() / ref.norm())
    assert err < tol, ("forward", err)

    G = torch.randn_like(ref)
    img.add_grad(G.float())
    tape.backward()
    torch.cuda.synchronize()
    (ref * G.float().double()).sum().backward()

    def rel(a, bb):
        return float((a.double() - bb).norm() / (bb.norm() + 1e-30))
    btol = tol * 3
    if h >= 128 and precision == "fp32x3":
        # millions of LeakyReLU units: a handful of pre-activations lie within the fp32x3 rounding distance (1e-5) of zero and
        # flip against the fp64 reference; each flip moves these relative L2 errors by ~1e-4 (tests/test_gpu_step.py docstring)
        btol = 3e-3
    gw2, gb2 = layer2.grad_views()
    gw, gb = layer.grad_views()
    errs = {"wgrad window-out": rel(gw2, w264.grad), "bias grad final": rel(gb2, b264.grad),
            "wgrad window-in": rel(gw, w64.grad), "bias grad": rel(gb, bb64.grad),
            "image grad": rel(ib.grad, b64i.grad)}
    print("\n[final conv / image io %s] %s" % (precision, errs))
    assert ia.grad is None
    # the first conv's quantities sit below a LeakyReLU: one flipped unit of this 8x8 plane is ~6e-3 (see test_gpu_step)
    lim = {k: (btol if "final" in k or "window-out" in k else max(btol, 2e-2)) for k in errs}
    assert all(errs[k] < lim[k] for k in errs), errs


UP_CASES = [
    # cin, cout, n, h (source plane h x h -> output 2h x 2h), out_pad
    (256, 128, 2, 64, 1),        # first up block of the 256x256 decoder (flattened-grid tiles, N = 512 in two tiles)
    (128, 64, 1, 128, 3),        # second up block: 128-pixel row tiles, N = 256
    (64, 32, 2, 16, 1),          # narrow ('tiny') networks: phases narrower than a staged row segment -> generic epilogue
    (32, 16, 2, 32, 3),
    (64, 64, 1, 8, 2),
    (128, 64, 2, 5, 1),          # odd, small plane: every source pixel within 2 of the border
]


@pytest.mark.parametrize("precision", ["fp32x3", "bf16"])
@pytest.mark.parametrize("cin,cout,n,h,out_pad", UP_CASES)
def test_up_block_subpixel(precision, cin, cout, n, h, out_pad):
    """nearest 2x upsample -> ReflectionPad2d(2) -> Conv2d 5x5 -> LayerNorm -> ReLU (reference networks.py:256-257,520-536) in
    sub-pixel form (engine.conv_block_up: 3x3 / 4*Cout main convolution + exact ring strips) vs the plain fp64 composition,
    forward and every gradient; includes the reflect-in-up-sampled-coordinates border ring."""
    eng = E.Engine(precision)
    tol = 2e-4 if precision == "fp32x3" else 4e-2
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(n, cin, h, h, device=dev)
    w = torch.nn.Parameter(torch.randn(cout, cin, 5, 5, device=dev) / (cin * 25) ** 0.5)
    b = torch.nn.Parameter(torch.randn(cout, device=dev) * 0.1)
    gamma, beta = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)
    arena = E.GradArena(eng.device)
    layer = E.ConvLayer(eng, arena, w, b, 1, 2)
    ln_off = (arena.reserve(cout), arena.reserve(cout))
    arena.finalize()
    up = E.UpConvLayer(eng, layer)
    xa = plane_from_nchw(eng, x, 1)
    xa.requires_grad = True
    ln = (gamma, beta, arena.view(ln_off[0], cout), arena.view(ln_off[1], cout))
    tape = E.Tape()
    out = eng.conv_block_up(tape, up, xa, act=N.ACT_RELU, out_pad=out_pad, ln=ln)
    torch.cuda.synchronize()

    def eff(t, planes):
        hi = t.detach().float().bfloat16()
        return (hi.double() + ((t.detach().float() - hi.float()).bfloat16().double() if planes == 2 else 0))
    P = eng.prec.planes
    x64 = eff(x, P).requires_grad_(True)
    w64 = eff(w, P).requires_grad_(True)
    b64 = b.detach().double().requires_grad_(True)
    g64, be64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    u = F.interpolate(x64, scale_factor=2, mode="nearest")
    ref = ref_block(u, w64, b64, 1, 2, N.NORM_LN, N.ACT_RELU, None, 1, out_pad, None, (g64, be64))
    p = out.pad
    full = out.buf[:, :out.numel].float().sum(0).view(n, out.h + 2 * p, out.w + 2 * p, out.c)
    got = full[..., :cout].permute(0, 3, 1, 2).double()
    assert got.shape == ref.shape
    err = float((got - ref).norm() / ref.norm())
    ring = torch.ones_like(ref, dtype=torch.bool)
    ring[:, :, p + 2:ref.shape[2] - p - 2, p + 2:ref.shape[3] - p - 2] = False
    err_ring = float(((got - ref) * ring).norm() / (ref * ring).norm())
    assert err < tol and err_ring < tol, ("forward", err, "ring", err_ring)
    if out.c > cout:
        assert float(full[..., cout:].abs().max()) == 0.0

    G = torch.randn_like(ref)
    gp = torch.zeros((n, out.h + 2 * p, out.w + 2 * p, out.c), device=dev, dtype=eng.prec.dtype)
    gp[..., :cout] = G.permute(0, 2, 3, 1).to(eng.prec.dtype)
    Geff = gp[..., :cout].double().permute(0, 3, 1, 2)
    out.gp = gp
    tape.backward()
    torch.cuda.synchronize()
    (ref * Geff).sum().backward()

    def rel(a, bb):
        return float((a.double() - bb).norm() / (bb.norm() + 1e-30))
    gx = xa.gp[..., :cin].double().permute(0, 3, 1, 2)
    probe = torch.zeros_like(x64).requires_grad_(True)
    (F.pad(probe, (1,) * 4, mode="reflect") * gx).sum().backward()
    btol = tol * 3
    if h >= 64 and precision == "fp32x3":
        btol = 3e-3          # ReLU units within the fp32x3 rounding distance of zero flip against fp64 (test_conv_block)
    gw, gb = layer.grad_views()
    errs = {"dgrad": rel(probe.grad, x64.grad), "wgrad": rel(gw, w64.grad), "bgrad": rel(gb, b64.grad),
            "dgamma": rel(ln[2], g64.grad), "dbeta": rel(ln[3], be64.grad)}
    print("\n[up block %d->%d %dx%d n=%d %s] forward %.2e (ring %.2e)  %s" % (
        cin, cout, h, h, n, precision, err, err_ring, "  ".join("%s %.2e" % kv for kv in errs.items())))
    bad = {k: v for k, v in errs.items() if not v < btol}
    assert not bad, bad
