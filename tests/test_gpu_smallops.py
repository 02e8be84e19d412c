"""-m gpu: the small operators of csrc/smallops.cu through the C ABI vs the torch ops the reference calls.
Index / mask ops are compared bit-exactly (AvgPool valid-count divisor vs F.avg_pool2d and its ATen backward, the focus blend
vs the reference's fp32 expression); reductions and small GEMMs against fp64 torch."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import aclgan_native as N
import gpu_util as G

pytestmark = pytest.mark.gpu
SP = G.stream_ptr


@pytest.mark.parametrize("n,c,h,w", [(2, 3, 256, 256), (3, 6, 128, 128), (1, 3, 7, 9), (2, 6, 5, 4), (1, 1, 1, 1), (1, 2, 2, 3)])
def test_avgpool_bit_exact(n, c, h, w):
    """nn.AvgPool2d(3, stride 2, padding 1, count_include_pad=False) (reference networks.py:33): divisor 4 | 6 | 9 by position"""
    torch.manual_seed(0)
    x = torch.randn(n, c, h, w, device="cuda")
    ref = F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)
    out = torch.empty_like(ref)
    a = N.AvgPoolArgs(x.data_ptr(), out.data_ptr(), n * c, h, w, 0)
    N.check(N.lib().aclgan_avgpool3x3s2_fwd(C.byref(a), SP()), "avgpool_fwd")
    assert torch.equal(out, ref)
    g = torch.randn_like(ref)
    gref = torch.ops.aten.avg_pool2d_backward(g, x, [3, 3], [2, 2], [1, 1], False, False, None)
    gx = torch.empty_like(x)
    b = N.AvgPoolArgs(g.data_ptr(), gx.data_ptr(), n * c, h, w, 0)
    N.check(N.lib().aclgan_avgpool3x3s2_bwd(C.byref(b), SP()), "avgpool_bwd")
    assert torch.equal(gx, gref)
    base = torch.randn_like(x)
    acc = base.clone()
    b.dst, b.accumulate = acc.data_ptr(), 1
    N.check(N.lib().aclgan_avgpool3x3s2_bwd(C.byref(b), SP()), "avgpool_bwd(acc)")
    assert torch.equal(acc, base + gref)


@pytest.mark.parametrize("planes,kind", [(1, 0), (2, 1)])
def test_style_head(planes, kind):
    """AdaptiveAvgPool2d(1) + Conv2d(C, 8, 1) (networks.py:222-223) forward / backward"""
    torch.manual_seed(1)
    n, c, h, w, sd = 3, 256, 4, 4, 8
    x = torch.randn(n, c, h, w, device="cuda")
    act, buf, xeff = G.make_act(x, 0, c, planes)
    W = torch.randn(sd, c, device="cuda") * 0.1
    b = torch.randn(sd, device="cuda")
    pooled = torch.empty(n, c, device="cuda")
    st = torch.empty(n, sd, device="cuda")
    a = N.StyleHeadArgs()
    a.x, a.c_valid, a.style_dim = act, c, sd
    a.weight, a.bias, a.pooled, a.style = W.data_ptr(), b.data_ptr(), pooled.data_ptr(), st.data_ptr()
    N.check(N.lib().aclgan_style_head_fwd(C.byref(a), SP()), "style_head_fwd")
    x64 = xeff.clone().requires_grad_(True)
    W64, b64 = W.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = F.conv2d(x64.mean(dim=(2, 3), keepdim=True), W64.view(sd, c, 1, 1), b64).view(n, sd)
    assert G.rel_err(st, ref) < 1e-5
    ds = torch.randn(n, sd, device="cuda")
    (ref * ds.double()).sum().backward()
    dW, db = torch.zeros_like(W), torch.zeros_like(b)
    gr = torch.empty(n, h, w, c, device="cuda", dtype=torch.float32 if kind else torch.bfloat16)
    a.dstyle, a.dweight, a.dbias, a.gr, a.g_kind = ds.data_ptr(), dW.data_ptr(), db.data_ptr(), gr.data_ptr(), kind
    N.check(N.lib().aclgan_style_head_bwd(C.byref(a), SP()), "style_head_bwd")
    assert G.rel_err(dW, W64.grad) < 1e-5 and G.rel_err(db, b64.grad) < 1e-5
    assert G.rel_err(gr.permute(0, 3, 1, 2).float(), x64.grad) < (1e-5 if kind else 5e-3)


@pytest.mark.parametrize("n", [2, 16, 19])
def test_mlp(n):
    """MLP 8 -> 256 -> 256 -> 4096 (networks.py:280-292) forward / backward vs torch autograd in fp64"""
    torch.manual_seed(2)
    dims = [8, 256, 256, 4096]
    Ws = [torch.randn(dims[i + 1], dims[i], device="cuda") * (1.0 / dims[i] ** 0.5) for i in range(3)]
    bs = [torch.randn(dims[i + 1], device="cuda") * 0.1 for i in range(3)]
    x = torch.randn(n, 8, device="cuda")
    hs = [x] + [torch.empty(n, d, device="cuda") for d in dims[1:]]
    a = N.MlpArgs()
    a.n, a.n_layers = n, 3
    for i in range(4):
        a.dims[i] = dims[i]
        a.h[i] = hs[i].data_ptr()
    for i in range(3):
        a.w[i], a.b[i] = Ws[i].data_ptr(), bs[i].data_ptr()
    a.h0_stride = 8
    N.check(N.lib().aclgan_mlp_fwd(C.byref(a), SP()), "mlp_fwd")
    W64 = [w.double().requires_grad_(True) for w in Ws]
    b64 = [b.double().requires_grad_(True) for b in bs]
    x64 = x.double().requires_grad_(True)
    h = torch.relu(F.linear(x64, W64[0], b64[0]))
    h = torch.relu(F.linear(h, W64[1], b64[1]))
    ref = F.linear(h, W64[2], b64[2])
    assert G.rel_err(hs[3], ref) < 1e-5
    g = torch.randn(n, 4096, device="cuda")
    (ref * g.double()).sum().backward()
    dWs = [torch.randn_like(w) for w in Ws]          # accumulated into (+=): start from non-zero contents
    dbs = [torch.randn_like(b) for b in bs]
    dW0 = [t.clone() for t in dWs]
    db0 = [t.clone() for t in dbs]
    scratch = [torch.empty(n, 256, device="cuda") for _ in range(2)]
    dx = torch.empty(n, 8, device="cuda")
    a.dh[3], a.dh[2], a.dh[1], a.dh[0] = g.data_ptr(), scratch[1].data_ptr(), scratch[0].data_ptr(), dx.data_ptr()
    for i in range(3):
        a.dw[i], a.db[i] = dWs[i].data_ptr(), dbs[i].data_ptr()
    N.check(N.lib().aclgan_mlp_bwd(C.byref(a), SP()), "mlp_bwd")
    for i in range(3):
        assert G.rel_err(dWs[i] - dW0[i], W64[i].grad) < 2e-5, i
        assert G.rel_err(dbs[i] - db0[i], b64[i].grad) < 2e-5, i
    assert G.rel_err(dx, x64.grad) < 2e-5


@pytest.mark.parametrize("scale,seed", [(0.05, 4), (0.4, 15)])
def test_dis_head_nsgan(scale, seed):
    """gan_kind NSGAN: F.binary_cross_entropy(F.sigmoid(o), t) per image group and its gradient seed (reference
    networks.py:68-72, 84-86, 99-103), against the fp32 ATen expression the reference evaluates (same clamps) and its autograd.
    The second case (CPU-seeded so it is the same everywhere) has one target-0 logit at +26: the fp32 sigmoid is exactly 1 there
    and the term is the -100 clamp of the logarithm; it has no target-0 logit in (11, 19), where 1 - sigmoid(o) is a handful of
    fp32 ulps and a last-bit difference of expf would move the term by O(1)."""
    torch.manual_seed(seed)
    groups, n_per, c, h, w = 2, 3, 512, 4, 4
    nimg = n_per * groups
    x = torch.randn(nimg, c, h, w).cuda()
    act, buf, xeff = G.make_act(x, 0, c, 2)
    W = (torch.randn(c) * scale).cuda()
    b = torch.randn(1).cuda()
    targets, weights = [0.0, 1.0], [1.0, 0.2]
    logits = torch.empty(nimg, 1, h, w, device="cuda")
    dl = torch.empty_like(logits)
    acc = torch.zeros(8, dtype=torch.float64, device="cuda")
    a = N.DisHeadArgs()
    a.x, a.c_valid, a.groups, a.gan_kind = act, c, groups, N.GAN_NSGAN
    a.weight, a.bias, a.logits, a.dlogits, a.loss = W.data_ptr(), b.data_ptr(), logits.data_ptr(), dl.data_ptr(), acc.data_ptr()
    for i in range(groups):
        a.target[i], a.gweight[i], a.loss_slot[i] = targets[i], weights[i], i
    N.check(N.lib().aclgan_dis_head_fwd(C.byref(a), SP()), "dis_head_fwd")
    g0 = logits[:n_per].flatten()
    assert int(((g0 > 11.0) & (g0 < 19.0)).sum()) == 0
    o = logits.detach().clone().requires_grad_(True)           # the loss as a function of the kernel's own fp32 logits
    total = 0
    for i in range(groups):
        og = o[i * n_per:(i + 1) * n_per]
        term = F.binary_cross_entropy(torch.sigmoid(og), torch.full_like(og, targets[i]))
        assert abs(float(acc[i]) - float(term)) < 2e-6 * float(term) + 1e-9, (i, float(acc[i]), float(term))
        total = total + weights[i] * term
    total.backward()
    assert bool(torch.isfinite(dl).all())
    assert G.rel_err(dl, o.grad) < 1e-5, G.rel_err(dl, o.grad)
    if scale > 0.1:
        assert float(g0.max()) > 20.0 and float(acc[0]) > 100.0 / g0.numel()      # the clamped term is in the mean
    a.gan_kind = 7
    assert N.lib().aclgan_dis_head_fwd(C.byref(a), SP()) == -3          # ACLGAN_ERR_UNSUPPORTED


@pytest.mark.parametrize("planes,kind,groups", [(1, 0, 3), (2, 1, 2), (1, 0, 1)])
def test_dis_head_lsgan(planes, kind, groups):
    """Conv2d(512, 1, 1) head (networks.py:45) + LSGAN terms / gradient seed (networks.py:67,83,98) per image group"""
    torch.manual_seed(3)
    n_per, c, h, w = 2, 512, 4, 4
    nimg = n_per * groups
    x = torch.randn(nimg, c, h, w, device="cuda")
    act, buf, xeff = G.make_act(x, 0, c, planes)
    W = torch.randn(c, device="cuda") * 0.05
    b = torch.randn(1, device="cuda")
    targets, weights = [1.0, 0.0, 0.0][:groups], [1.0, 0.5, 0.25][:groups]
    logits = torch.empty(nimg, 1, h, w, device="cuda")
    dl = torch.empty_like(logits)
    acc = torch.zeros(8, dtype=torch.float64, device="cuda")
    a = N.DisHeadArgs()
    a.x, a.c_valid, a.groups = act, c, groups
    a.weight, a.bias, a.logits, a.dlogits, a.loss = W.data_ptr(), b.data_ptr(), logits.data_ptr(), dl.data_ptr(), acc.data_ptr()
    for i in range(groups):
        a.target[i], a.gweight[i], a.loss_slot[i] = targets[i], weights[i], 2 * i + 1
    N.check(N.lib().aclgan_dis_head_fwd(C.byref(a), SP()), "dis_head_fwd")
    x64 = xeff.clone().requires_grad_(True)
    W64, b64 = W.double().requires_grad_(True), b.double().requires_grad_(True)
    o = F.conv2d(x64, W64.view(1, c, 1, 1), b64)
    assert G.rel_err(logits, o) < 1e-5
    total = 0
    for i in range(groups):
        term = torch.mean((o[i * n_per:(i + 1) * n_per] - targets[i]) ** 2)
        assert abs(float(acc[2 * i + 1]) - float(term)) < 1e-5 * float(term) + 1e-9, i
        total = total + weights[i] * term
    o.retain_grad()
    total.backward()
    assert G.rel_err(dl, o.grad) < 1e-5
    dW, db = torch.zeros_like(W), torch.zeros_like(b)
    gr = torch.empty(nimg, h, w, c, device="cuda", dtype=torch.float32 if kind else torch.bfloat16)
    bb = N.DisHeadBwdArgs()
    bb.x, bb.c_valid, bb.g_kind = act, c, kind
    bb.weight, bb.dlogits, bb.dweight, bb.dbias, bb.gr = W.data_ptr(), dl.data_ptr(), dW.data_ptr(), db.data_ptr(), gr.data_ptr()
    N.check(N.lib().aclgan_dis_head_bwd(C.byref(bb), SP()), "dis_head_bwd")
    assert G.rel_err(dW, W64.grad) < 2e-5 and G.rel_err(db, b64.grad) < 2e-5
    assert G.rel_err(gr.permute(0, 3, 1, 2).float(), x64.grad) < (1e-5 if kind else 5e-3)


def test_focus_blend_bit_exact_and_adjoint():
    """focus_translation (trainer.py:85-88): forward bit-identical to the reference's fp32 expression; backward vs autograd"""
    torch.manual_seed(4)
    n, h, w = 2, 32, 48
    out4 = torch.tanh(torch.randn(n, 4, h, w, device="cuda"))
    bg = torch.rand(n, 3, h, w, device="cuda") * 2 - 1
    dst = torch.empty(n, 3, h, w, device="cuda")
    a = N.BlendArgs()
    a.out4, a.bg, a.dst, a.n, a.h, a.w = out4.data_ptr(), bg.data_ptr(), dst.data_ptr(), n, h, w
    N.check(N.lib().aclgan_focus_blend_fwd(C.byref(a), SP()), "blend_fwd")
    fg, focus = out4.split(3, 1)
    x_map = ((focus + 1) / 2).repeat(1, 3, 1, 1)
    assert torch.equal(dst, fg * x_map + bg * (1 - x_map))
    o64, b64 = out4.double().requires_grad_(True), bg.double().requires_grad_(True)
    m = ((o64[:, 3:4] + 1) / 2)
    ref = o64[:, :3] * m + b64 * (1 - m)
    d = torch.randn(n, 3, h, w, device="cuda")
    (ref * d.double()).sum().backward()
    do = torch.randn(n, 4, h, w, device="cuda")
    do0 = do.clone()
    dbg = torch.empty(n, 3, h, w, device="cuda")
    a.ddst, a.dout4, a.dbg, a.acc_out4, a.acc_bg = d.data_ptr(), do.data_ptr(), dbg.data_ptr(), 1, 0
    N.check(N.lib().aclgan_focus_blend_bwd(C.byref(a), SP()), "blend_bwd")
    assert G.rel_err(do - do0, o64.grad) < 1e-5 and G.rel_err(dbg, b64.grad) < 1e-6


def test_l1_and_focus_losses():
    """recon_criterion (trainer.py:61-62) and the focus size / digit losses (trainer.py:146-161) incl. their gradients"""
    torch.manual_seed(5)
    n, h, w = 3, 64, 64
    out4 = torch.tanh(torch.randn(n, 4, h, w, device="cuda") * 0.7)
    tgt = torch.rand(n, 3, h, w, device="cuda") * 2 - 1
    acc = torch.zeros(24, dtype=torch.float64, device="cuda")
    # ---- L1 on the first 3 of 4 channels, gradient accumulated into a [n,4,h,w] buffer
    rw = 10.0
    da = torch.zeros(n, 4, h, w, device="cuda")
    a = N.LossReduceArgs()
    a.mode, a.n, a.ca, a.c, a.h, a.w = N.LOSS_L1, n, 4, 3, h, w
    a.a, a.b, a.acc, a.slot = out4.data_ptr(), tgt.data_ptr(), acc.data_ptr(), 17
    a.da, a.acc_da, a.gscale = da.data_ptr(), 1, rw / (n * 3 * h * w)
    N.check(N.lib().aclgan_loss_reduce(C.byref(a), SP()), "l1")
    o64 = out4.double().requires_grad_(True)
    l1 = torch.mean(torch.abs(o64[:, :3] - tgt.double()))
    (rw * l1).backward()
    assert abs(float(acc[17]) - float(l1)) < 1e-6 * float(l1)
    assert G.rel_err(da, o64.grad) < 1e-6 and float(da[:, 3].abs().max()) == 0.0
    # ---- focus losses on channel 3: two regimes (mask mean above `upper` -> size loss active; between the bounds -> only digit)
    for shift in (-0.3, 0.9):
        o4 = (out4 + torch.tensor([0, 0, 0, shift], device="cuda").view(1, 4, 1, 1)).clamp(-1, 1).contiguous()
        acc.zero_()
        up, lo, delta, eps, lam = 0.5, 0.3, 0.001, 0.01, 0.025
        f = N.LossReduceArgs()
        f.mode, f.n, f.ca, f.c, f.h, f.w = N.LOSS_FOCUS, n, 4, 1, h, w
        f.a, f.acc, f.slot, f.upper, f.lower, f.eps = o4.data_ptr(), acc.data_ptr(), 5, up, lo, eps
        N.check(N.lib().aclgan_loss_reduce(C.byref(f), SP()), "focus")
        dout = torch.zeros(n, 4, h, w, device="cuda")
        gscale = 0.5 * lam / (h * w * n * 3)
        g = N.FocusGradArgs()
        g.out4, g.dout4, g.n, g.h, g.w, g.slot, g.size_slot, g.acc = o4.data_ptr(), dout.data_ptr(), n, h, w, 5, 8, 1
        g.sums, g.delta, g.eps, g.gscale = acc.data_ptr(), delta, eps, gscale
        N.check(N.lib().aclgan_focus_grad(C.byref(g), SP()), "focus_grad")
        o64 = o4.double().requires_grad_(True)
        mm = (o64[:, 3:4] + 1) / 2
        size = torch.relu(torch.sum(mm - up)) ** 2 * delta + torch.relu(torch.sum(lo - mm)) ** 2 * delta
        digit = torch.sum(1 / (torch.abs(mm - 0.5) + eps))
        (lam * (size + digit) / h / w / n / 3).backward()
        assert abs(float(acc[7]) - float(digit)) < 1e-5 * float(digit)
        assert abs(float(acc[8]) - float(size)) <= 1e-4 * float(size) + 1e-9, (float(acc[8]), float(size))
        assert (float(size) > 0) == (shift > 0)
        assert G.rel_err(dout[:, 3], o64.grad[:, 3]) < 1e-4 and float(dout[:, :3].abs().max()) == 0.0
    # ---- combination
    M = torch.randn(5, 24, device="cuda")
    out = torch.empty(5, device="cuda")
    N.check(N.lib().aclgan_loss_combine(acc.data_ptr(), M.data_ptr(), out.data_ptr(), 5, 24, SP()), "combine")
    assert G.rel_err(out, M.double() @ acc) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_axpby_zero_copy(dtype):
    torch.manual_seed(6)
    a = torch.randn(100003, device="cuda").to(dtype)
    b = torch.randn(100003, device="cuda").to(dtype)
    d = torch.empty_like(a)
    kind = 1 if dtype == torch.float32 else 0
    N.check(N.lib().aclgan_axpby(d.data_ptr(), a.data_ptr(), b.data_ptr(), 0.5, 2.0, a.numel(), kind, SP()), "axpby")
    assert torch.equal(d, (0.5 * a.float() + 2.0 * b.float()).to(dtype))
    N.check(N.lib().aclgan_axpby(a.data_ptr(), a.data_ptr(), b.data_ptr(), 1.0, 1.0, a.numel(), kind, SP()), "axpby(in place)")
    torch.cuda.synchronize()
    c = torch.empty_like(a)
    assert N.lib().aclgan_copy(c.data_ptr(), a.data_ptr(), a.numel() * a.element_size(), SP()) == 0
    assert torch.equal(c, a)
    assert N.lib().aclgan_zero(c.data_ptr(), c.numel() * c.element_size(), SP()) == 0
    assert float(c.float().abs().max()) == 0.0


@pytest.mark.parametrize("cfgname", ["male2female.yaml", "selfie2anime.yaml"])
def test_step_pair_launches_no_library_kernels(cfgname):
    """every kernel of a graph-replayed step-pair (bf16, the benchmarked mode) is one of
    libaclgan_b200.so's - no ATen element-wise / reduce / cat / fill kernels and no cuBLAS GEMMs (memset / memcpy nodes are not
    kernels).  Kernel names come from CUPTI through torch.profiler."""
    import copy
    import os
    import yaml
    import trainer as T
    from torch.profiler import ProfilerActivity, profile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, "acl-gan_b200", "configs", cfgname)))
    cfg["gen"].update(dim=16, mlp_dim=32, n_res=2)
    cfg["dis"].update(dim=16)
    cfg["display_size"] = 2
    cfg["precision"] = "bf16"
    cfg["expose_grads"] = 0            # (re-materialising the flipped .grad view of the final conv is an ATen copy outside the graph)
    torch.manual_seed(0)
    tr = T.aclgan_Trainer(copy.deepcopy(cfg)).cuda()
    xa = torch.rand(2, 3, 64, 64, device="cuda") * 2 - 1
    xb = torch.rand(2, 3, 64, 64, device="cuda") * 2 - 1
    for _ in range(2):
        tr.dis_update(xa, xb, cfg)
        tr.gen_update(xa, xb, cfg)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.dis_update(xa, xb, cfg)
        tr.gen_update(xa, xb, cfg)
        torch.cuda.synchronize()
    names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    kernels = [n for n in names if not n.lower().startswith(("memcpy", "memset"))]
    if not kernels:
        pytest.skip("CUPTI kernel records unavailable on this box")
    foreign = sorted(set(n.split("(")[0][:80] for n in kernels if "aclgan::" not in n))
    print("\n[launch list %s] %d kernel launches per step-pair, %d of libaclgan_b200.so; others: %s" % (
        cfgname, len(kernels), sum("aclgan::" in n for n in kernels), foreign))
    assert not foreign, foreign
