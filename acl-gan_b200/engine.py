"""Execution engine of the B200-native ACL-GAN step: activation planes, per-conv packed state, the fused
block operators (forward + hand-scheduled backward on a tape) that networks.py / trainer.py compose.

Nothing here falls back to eager PyTorch convolutions or to the CPU: every heavy operator is a launch of
libaclgan_b200.so through the C ABI (include/aclgan_b200.h).  PyTorch provides device memory, streams and the
tiny glue on [N, C]-sized vectors.

Reference behaviour being reproduced (file:line under /root/reference): Conv2dBlock.forward networks.py:365-371,
ResBlock.forward :306-310, nn.Upsample :256, AdaptiveInstanceNorm2d.forward :490-503, LayerNorm.forward
:520-536, and autograd's backward of all of them under loss.backward() (trainer.py:169,292).
"""
import contextlib
import ctypes as C
import math
import os

import torch

import aclgan_native as N


def _sp():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else 0)


def round_up(a, b):
    return (a + b - 1) // b * b


def _rc(rc, what):
    """status check of a C-ABI call that launches no kernel (memset / memcpy nodes)"""
    if rc != 0:
        raise N.NativeError("%s failed with status %d" % (what, rc))


def zero_(t):
    """cudaMemsetAsync on the current stream (a memset node under graph capture) instead of an ATen fill kernel"""
    if not t.is_cuda:           # host-side bookkeeping objects in the CPU unit tests (the compute path is CUDA only)
        return t.zero_()
    _rc(N.lib().aclgan_zero(t.data_ptr(), t.numel() * t.element_size(), _sp()), "zero")
    return t


def zeros(shape, dtype, device):
    return zero_(torch.empty(shape, dtype=dtype, device=device))


@contextlib.contextmanager
def _plan_env(name, value):
    """sets a plan-time triage switch (csrc/plans.cu reads them with getenv) for the duration of one host-side plan call"""
    if value is None:
        yield
        return
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        yield
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


def copy_(dst, src):
    """contiguous device-to-device copy as a memcpy node"""
    assert dst.is_contiguous() and src.is_contiguous() and dst.dtype == src.dtype and dst.numel() == src.numel()
    _rc(N.lib().aclgan_copy(dst.data_ptr(), src.data_ptr(), dst.numel() * dst.element_size(), _sp()), "copy")
    return dst


def axpby(dst, a, b, alpha=1.0, beta=1.0):
    """dst = alpha * a + beta * b (b may be None) on bf16 / fp32 contiguous tensors: gradient accumulation without ATen"""
    assert dst.is_contiguous() and a.is_contiguous() and dst.dtype == a.dtype and dst.numel() == a.numel()
    assert dst.dtype in (torch.float32, torch.bfloat16)
    if b is not None:
        assert b.is_contiguous() and b.dtype == a.dtype and b.numel() == a.numel()
    N.check(N.lib().aclgan_axpby(dst.data_ptr(), a.data_ptr(), b.data_ptr() if b is not None else 0, float(alpha), float(beta),
                                 dst.numel(), 1 if dst.dtype == torch.float32 else 0, _sp()), "axpby")
    return dst


def acc_(dst, src):
    """dst += src"""
    return axpby(dst, dst, src, 1.0, 1.0)


def cat0(tensors):
    """torch.cat along dim 0 of contiguous tensors as memcpy nodes"""
    if len(tensors) == 1:
        return tensors[0]
    out = torch.empty((sum(t.shape[0] for t in tensors),) + tuple(tensors[0].shape[1:]), dtype=tensors[0].dtype,
                      device=tensors[0].device)
    off = 0
    for t in tensors:
        copy_(out[off:off + t.shape[0]], t.contiguous())
        off += t.shape[0]
    return out


class Precision:
    """bf16: one bf16 plane, bf16 raw conv outputs and gradients (throughput mode).
    fp32x3: hi/lo bf16 planes (3 tensor-core passes ~ fp32 products), fp32 raw outputs and gradients (parity mode)."""

    def __init__(self, name):
        assert name in ("bf16", "fp32x3"), name
        self.name = name
        self.planes = 1 if name == "bf16" else 2
        self.kind = 0 if name == "bf16" else 1          # dense tensor kind: 0 bf16, 1 fp32
        self.dtype = torch.bfloat16 if name == "bf16" else torch.float32


class Tape:
    """Hand-scheduled reverse pass: forward operators push closures, backward() runs them in reverse.

    Closures recorded while `tag` is set (trainer: one tag per discriminator pass) belong to an independent chain; in
    backward() consecutive chains run on their own side streams (fork from / join into the caller's stream), so the
    many sub-wave kernels of the small discriminator scales overlap instead of serialising.  Under CUDA-graph capture
    the fork / join become parallel branches of the graph."""

    def __init__(self, enabled=True):
        self.ops = []
        self.enabled = enabled
        self.tag = None

    def push(self, fn):
        if self.enabled:
            self.ops.append((fn, self.tag))

    def backward(self, streams=None):
        main = torch.cuda.current_stream() if (streams and torch.cuda.is_available()) else None
        cur, used = None, set()
        while self.ops:
            fn, tag = self.ops.pop()
            if main is None or tag is None:
                if main is not None and used:
                    for t in used:                      # join: everything after this sees the chains' results
                        main.wait_stream(streams[t])
                    used.clear()
                cur = None
                fn()
                continue
            if tag != cur:
                streams[tag].wait_stream(main)          # fork point = everything the main stream has issued so far
                used.add(tag)
                cur = tag
            with torch.cuda.stream(streams[tag]):
                fn()
        if main is not None:
            for t in used:
                main.wait_stream(streams[t])


# Debug switch (tools/poison_check.py, ACLGAN_POISON=1): planes that are allocated WITHOUT a memset are filled with bf16 quiet
# NaNs, so a halo / channel-pad element that a consumer reads but no producer wrote surfaces as a non-finite loss instead of
# depending on what the allocator handed out.  Off in the product path (the fill is an ATen kernel).
POISON = os.environ.get("ACLGAN_POISON", "0") != "0"


class ActT:
    """Reflect-padded NHWC activation plane(s): buf [planes, numel + slack] bf16."""

    def __init__(self, eng, n, h, w, c_valid, pad, cs=None, zero=False):
        self.eng = eng
        self.n, self.h, self.w, self.c_valid, self.pad = n, h, w, c_valid, pad
        self.c = cs if cs is not None else round_up(c_valid, 64)
        self.planes = eng.prec.planes
        self.numel = n * (h + 2 * pad) * (w + 2 * pad) * self.c
        # planes with fewer than 64 channels are read through pixel-window tensor maps that overrun the plane by up to
        # 64 elements (zeroed slack); full-width planes are only ever read inside their extent (TMA zero-fills beyond)
        self.buf = torch.empty((self.planes, self.numel + 64), dtype=torch.bfloat16, device=eng.device)
        if zero or self.c < 64 or self.c != c_valid:
            zero_(self.buf)
        elif POISON:
            self.buf.view(torch.int16).fill_(0x7FC0)
        self.gp = None            # gradient of the padded plane [n, h+2p, w+2p, c] (engine gradient dtype)
        self.gr = None            # dense gradient [n, h, w, c] (residual branches / heads)
        self.requires_grad = False

    def struct(self):
        a = N.Act()
        for p in range(self.planes):
            a.data[p] = self.buf[p].data_ptr()
        a.planes, a.n, a.h, a.w, a.c, a.pad = self.planes, self.n, self.h, self.w, self.c, self.pad
        return a

    def add_gp(self, g):
        self.gp = g if self.gp is None else acc_(self.gp, g)

    def add_gr(self, g):
        self.gr = g if self.gr is None else acc_(self.gr, g)

    def value_nchw(self):
        """fp32 NCHW copy of the logical (un-padded, valid-channel) content - for tests / API boundaries."""
        p = self.pad
        v = self.buf[:, :self.numel].float().sum(0).view(self.n, self.h + 2 * p, self.w + 2 * p, self.c)
        v = v[:, p:p + self.h, p:p + self.w, :self.c_valid]
        return v.permute(0, 3, 1, 2).contiguous()


class ImgT:
    """NCHW fp32 image tensor with an optional gradient accumulator (images cross the API boundary in the
    reference's own layout: trainer.py passes B x 3 x H x W fp32 tensors)."""

    def __init__(self, t, requires_grad=False):
        self.t = t.contiguous()
        self.requires_grad = requires_grad
        self.grad = None

    def add_grad(self, g):
        self.grad = g if self.grad is None else acc_(self.grad, g.contiguous())

    def grad_buffer(self):
        """(tensor to write d/d(self.t) into, accumulate flag): the existing gradient (accumulate) or a fresh one (assign)"""
        if self.grad is None:
            self.grad = torch.empty_like(self.t)
            return self.grad, 0
        return self.grad, 1

    def zero_grad_buffer(self):
        """gradient buffer that may be accumulated into channel-wise: zero-filled (memset) when it does not exist yet"""
        if self.grad is None:
            self.grad = zeros(self.t.shape, self.t.dtype, self.t.device)
        return self.grad


class SumsPool:
    """fp64 workspace for the per-(n, c) statistics of one update: ONE memset per update instead of one fill kernel per
    layer.  In counting mode (buf is None) it hands out fresh zeroed tensors and records the total size."""

    def __init__(self, device, numel=None):
        self.buf = None if numel is None else torch.zeros(max(2, numel), dtype=torch.float64, device=device)
        self.device = device
        self.off = 0

    def begin(self):
        self.off = 0
        if self.buf is not None:
            zero_(self.buf)

    def take(self, shape):
        numel = 1
        for d in shape:
            numel *= d
        numel = round_up(numel, 2)
        if self.buf is None or self.off + numel > self.buf.numel():
            self.off += numel
            return zeros(shape, torch.float64, self.device)
        t = self.buf[self.off:self.off + numel].view(shape)
        self.off += numel
        return t


class GradArena:
    """One flat fp32 buffer per optimizer group (generators / discriminators) holding every gradient of the
    group: a single memset per step and a single NCCL all-reduce."""

    def __init__(self, device):
        self.device = device
        self.sizes = []
        self.flat = None

    def reserve(self, numel):
        off = sum(self.sizes)
        self.sizes.append(round_up(numel, 4))       # keep 16 B alignment for float4 atomics
        return off

    def finalize(self):
        self.flat = torch.zeros(max(4, sum(self.sizes)), dtype=torch.float32, device=self.device)

    def view(self, off, numel):
        return self.flat[off:off + numel]

    def zero_(self):
        zero_(self.flat)


def _affine_index(desc, transposed, shape):
    """coefficients of the (affine) packed-weight index function of the C library"""
    L = N.lib()
    co, ci, kh, kw = shape

    def idx(a, b, c, d):
        return L.aclgan_packed_weight_index(C.byref(desc), int(transposed), a, b, c, d)

    base = idx(0, 0, 0, 0)
    return (base, idx(1, 0, 0, 0) - base if co > 1 else 0, idx(0, 1, 0, 0) - base if ci > 1 else 0,
            idx(0, 0, 1, 0) - base if kh > 1 else 0, idx(0, 0, 0, 1) - base if kw > 1 else 0)


class ConvLayer:
    """Engine-side state of one nn.Conv2d parameter pair: packed bf16 weights (forward / transposed, hi[/lo])
    derived from the fp32 OIHW master parameter, and the slice of the gradient arena its wgrad kernel fills."""

    def __init__(self, eng, arena, weight, bias, stride, pad, window=N.WINDOW_NONE):
        self.eng = eng
        self.weight, self.bias = weight, bias
        co, ci, k, _ = weight.shape
        self.cout, self.cin, self.k, self.stride, self.pad, self.window = co, ci, k, stride, pad, window
        self.desc = N.ConvDesc(ci, co, k, stride, pad, window)
        L = N.lib()
        self.packed = {}
        self.aff = {}
        for tr in (0, 1):
            rows, kt = C.c_int64(), C.c_int64()
            L.aclgan_packed_weight_shape(C.byref(self.desc), tr, C.byref(rows), C.byref(kt))
            self.packed[tr] = torch.zeros((eng.prec.planes, rows.value * kt.value), dtype=torch.bfloat16,
                                          device=eng.device)
            self.aff[tr] = _affine_index(self.desc, tr, weight.shape)
            setattr(self, "shape%d" % tr, (rows.value, kt.value))
        self.layout = L.aclgan_wgrad_layout(C.byref(self.desc))
        rows, kt = getattr(self, "shape%d" % self.layout)
        self.arena = arena
        self.dw_off = arena.reserve(rows * kt)
        self.dw_numel = rows * kt
        self.db_off = arena.reserve(co)
        self.dirty = True
        self.derived = []           # UpConvLayer views that re-derive their weights whenever the master changes

    # -- packed weights -------------------------------------------------------------------------------
    def repack(self):
        L = N.lib()
        w = self.weight.detach()
        co, ci, kh, kw = w.shape
        for tr in (0, 1):
            a = N.PackWeightArgs()
            a.w = w.data_ptr()
            a.co, a.ci, a.kh, a.kw = co, ci, kh, kw
            a.base, a.s_co, a.s_ci, a.s_kh, a.s_kw = self.aff[tr]
            for p in range(self.eng.prec.planes):
                a.dst[p] = self.packed[tr][p].data_ptr()
            a.planes = self.eng.prec.planes
            N.check(L.aclgan_pack_weight(C.byref(a), _sp()), "pack_weight")
        self.dirty = False
        for d in self.derived:
            d.derive()

    def wptr(self, tr):
        if self.dirty:
            self.repack()
        t = self.packed[tr]
        return (C.c_uint64 * 2)(t[0].data_ptr(), t[1].data_ptr() if t.shape[0] > 1 else 0)

    # -- gradients ------------------------------------------------------------------------------------
    def dw(self):
        return self.arena.view(self.dw_off, self.dw_numel)

    def db(self):
        return self.arena.view(self.db_off, self.cout)

    def grad_layout(self):
        """(element offset in the arena, (s_co, s_ci, s_kh, s_kw)) of the OIHW gradient inside the packed-layout buffer"""
        base, s_co, s_ci, s_kh, s_kw = self.aff[self.layout]
        return self.dw_off + base, (s_co, s_ci, s_kh, s_kw)

    def grad_views(self):
        """(weight grad in OIHW shape, bias grad) as views of the arena (a small copy for the one layout whose
        kw stride is negative)."""
        base, s_co, s_ci, s_kh, s_kw = self.aff[self.layout]
        co, ci, kh, kw = self.weight.shape
        dw = self.dw()
        off = dw.storage_offset() + base          # as_strided offsets are absolute within the arena storage
        if s_kw >= 0:
            g = torch.as_strided(dw, (co, ci, kh, kw), (s_co, s_ci, s_kh, s_kw), off)
        else:
            g = torch.as_strided(dw, (co, ci, kh, kw), (s_co, s_ci, s_kh, -s_kw), off + (kw - 1) * s_kw).flip(3)
        return g, self.db()


class _LayerView:
    """what conv_fwd_launch / conv_dgrad / conv_wgrad need from a layer, for derived weight sets"""

    def __init__(self, desc, cin, cout, k, wptr, dw):
        self.desc, self.cin, self.cout, self.k, self.stride = desc, cin, cout, k, 1
        self.wptr, self.dw = wptr, dw


class UpConvLayer:
    """Sub-pixel form of an up-block convolution (nearest 2x upsample -> ReflectionPad2d(2) -> Conv2d 5x5, reference
    networks.py:256-257): derived state around the 5x5 ConvLayer `base` - the phase weights of the 3x3 / 4*Cout main convolution
    (both packings), the transposed-filter packings for the column strips, the tiled bias and the phase-weight gradient scratch
    (csrc/upconv.cu; geometry: tools/subpixel_pipeline.py)."""

    def __init__(self, eng, base):
        assert base.k == 5 and base.stride == 1 and base.pad == 2 and base.cout % 8 == 0
        self.eng, self.base = eng, base
        co, ci = base.cout, base.cin
        L = N.lib()
        self.desc = N.ConvDesc(ci, 4 * co, 3, 1, 1, N.WINDOW_NONE)
        self.packed, self.aff = {}, {}
        for tr in (0, 1):
            rows, kt = C.c_int64(), C.c_int64()
            L.aclgan_packed_weight_shape(C.byref(self.desc), tr, C.byref(rows), C.byref(kt))
            self.packed[tr] = zeros((eng.prec.planes, rows.value * kt.value), torch.bfloat16, eng.device)
            self.aff[tr] = _affine_index(self.desc, tr, (4 * co, ci, 3, 3))
            setattr(self, "shape%d" % tr, (rows.value, kt.value))
        self.layout = L.aclgan_wgrad_layout(C.byref(self.desc))
        rows, kt = getattr(self, "shape%d" % self.layout)
        self.dwp = torch.empty(rows * kt, dtype=torch.float32, device=eng.device)
        self.bias4 = torch.empty(4 * co, dtype=torch.float32, device=eng.device)
        self.packedT = {tr: torch.zeros_like(base.packed[tr]) for tr in (0, 1)}
        self.main = _LayerView(self.desc, ci, 4 * co, 3, self.wptr, lambda: self.dwp)
        self.colview = _LayerView(base.desc, ci, co, 5, self.wptrT, base.dw)
        base.derived.append(self)
        if not base.dirty:
            self.derive()

    @staticmethod
    def _ptrs(t):
        return (C.c_uint64 * 2)(t[0].data_ptr(), t[1].data_ptr() if t.shape[0] > 1 else 0)

    def wptr(self, tr):
        if self.base.dirty:
            self.base.repack()
        return self._ptrs(self.packed[tr])

    def wptrT(self, tr):
        if self.base.dirty:
            self.base.repack()
        return self._ptrs(self.packedT[tr])

    def derive(self):
        """phase weights / tiled bias / transposed-filter packings from the fp32 master (after every optimizer step)"""
        L = N.lib()
        base = self.base
        w = base.weight.detach()
        planes = self.eng.prec.planes
        a = N.UpDeriveArgs()
        a.w5, a.bias, a.co, a.ci, a.planes = w.data_ptr(), base.bias.data_ptr(), base.cout, base.cin, planes
        for tr in (0, 1):
            for p in range(planes):
                a.pk[tr][p] = self.packed[tr][p].data_ptr()
            for j in range(5):
                a.aff[tr][j] = self.aff[tr][j]
        a.bias4 = self.bias4.data_ptr()
        N.check(L.aclgan_up_derive_weights(C.byref(a), _sp()), "up_derive_weights")
        co, ci, kh, kw = w.shape
        for tr in (0, 1):       # W5^T: element (kh, kw) goes where the base packing puts (kw, kh)
            b = N.PackWeightArgs()
            b.w, b.co, b.ci, b.kh, b.kw = w.data_ptr(), co, ci, kh, kw
            b.base, b.s_co, b.s_ci, b.s_kw, b.s_kh = base.aff[tr]
            for p in range(planes):
                b.dst[p] = self.packedT[tr][p].data_ptr()
            b.planes = planes
            N.check(L.aclgan_pack_weight(C.byref(b), _sp()), "pack_weight(T)")


class Engine:
    def __init__(self, precision="bf16", device="cuda"):
        self.prec = Precision(precision) if isinstance(precision, str) else precision
        self.device = torch.device(device)
        self._check_device()
        N.lib()
        self.eps = 1e-5
        self.debug = None           # tests: dict collecting per-layer backward intermediates
        self.pool = None            # SumsPool of the update being recorded (trainer), or None: per-layer zeroed tensors
        self.fuse_stats = os.environ.get("ACLGAN_FUSE_STATS", "1") != "0"
        self.fuse_finalize = os.environ.get("ACLGAN_FUSE_FINALIZE", "0") != "0"

    def sums(self, n, cs):
        if self.pool is not None:
            return self.pool.take((n, cs, 2))
        return zeros((n, cs, 2), torch.float64, self.device)

    def _check_device(self):
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise N.NativeError("aclgan_b200 has no CPU path: a CUDA device is required")

    # ------------------------------------------------------------------------------------------ helpers
    def new_dense(self, n, h, w, c, zero=False):
        t = torch.empty((n, h, w, c), dtype=self.prec.dtype, device=self.device)
        if POISON and not zero:
            t.fill_(float("nan"))
        return zero_(t) if zero else t

    def t4(self, t):
        n, h, w, c = t.shape
        return N.Tensor4(t.data_ptr(), self.prec.kind, n, h, w, c)

    # ------------------------------------------------------------------------------------------ image I/O
    def pack_image(self, tape, img0, pad, cs, img1=None):
        """NCHW fp32 image (optionally the channel concat of two images) -> reflect-padded plane."""
        t0 = img0.t
        n, c0, h, w = t0.shape
        c1 = img1.t.shape[1] if img1 is not None else 0
        out = ActT(self, n, h, w, c0 + c1, pad, cs=cs, zero=True)
        a = N.PackImgArgs()
        a.src0 = t0.data_ptr()
        a.src1 = img1.t.data_ptr() if img1 is not None else 0
        a.c0, a.c1, a.n, a.h, a.w = c0, c1, n, h, w
        a.dst = out.struct()
        N.check(N.lib().aclgan_pack_img(C.byref(a), _sp()), "pack_img")
        out.requires_grad = img0.requires_grad or (img1 is not None and img1.requires_grad)
        if out.requires_grad:
            def bwd():
                if out.gp is None:
                    return
                for img, off in ((img0, 0), (img1, c0)):
                    if img is None or not img.requires_grad:
                        continue
                    u = N.ImgGradUnpackArgs()
                    g, acc = img.grad_buffer()
                    u.src = out.gp.data_ptr()
                    u.n, u.c, u.h, u.w = n, img.t.shape[1], h, w
                    u.cs, u.pad, u.c_off = out.gp.shape[-1], pad, off
                    u.dst, u.accumulate = g.data_ptr(), acc
                    N.check(N.lib().aclgan_img_grad_unpack(C.byref(u), _sp()), "img_grad_unpack")
                out.gp = None
            tape.push(bwd)
        return out

    def dup_plane(self, tape, x):
        """two copies of a plane along the batch axis (one content code decoded with two styles in ONE batched pass:
        trainer.py:109,113 decode c_2 twice); the gradients of both halves are summed back into `x`"""
        out = ActT(self, 2 * x.n, x.h, x.w, x.c_valid, x.pad)
        for p in range(x.planes):
            copy_(out.buf[p, :x.numel], x.buf[p, :x.numel])
            copy_(out.buf[p, x.numel:2 * x.numel], x.buf[p, :x.numel])
        out.requires_grad = x.requires_grad
        if tape.enabled and x.requires_grad:
            def bwd():
                n = x.n
                if out.gp is not None:
                    x.add_gp(axpby(torch.empty_like(out.gp[:n]), out.gp[:n], out.gp[n:]))
                if out.gr is not None:
                    x.add_gr(axpby(torch.empty_like(out.gr[:n]), out.gr[:n], out.gr[n:]))
                out.gp = out.gr = None
            tape.push(bwd)
        return out

    # ------------------------------------------------------------------------------------------ conv pieces
    def _out_plane(self, dst, act, bias, slope=0.2):
        """OutSpec that makes a conv epilogue write straight into plane `dst` (interior + reflect halo)."""
        o = N.OutSpec()
        hp, wp = dst.h + 2 * dst.pad, dst.w + 2 * dst.pad
        for p in range(dst.planes):
            o.ptr[p] = dst.buf[p].data_ptr()
        o.kind = N.OUT_BF16 if dst.planes == 1 else N.OUT_SPLIT
        o.act, o.slope, o.mirror = act, slope, dst.pad
        o.off = (dst.pad * wp + dst.pad) * dst.c
        o.sn, o.sy, o.sx, o.sc = hp * wp * dst.c, wp * dst.c, dst.c, 1
        o.N, o.H, o.W, o.C = dst.n, dst.h, dst.w, dst.c
        o.bias = bias.data_ptr() if bias is not None else 0
        o.bias_n = bias.numel() if bias is not None else 0
        return o

    def _out_dense(self, t, bias=None, act=N.ACT_NONE):
        n, h, w, c = t.shape
        o = N.OutSpec()
        o.ptr[0] = t.data_ptr()
        o.kind = N.OUT_BF16 if t.dtype == torch.bfloat16 else N.OUT_F32
        o.act, o.slope, o.mirror, o.off = act, 0.2, 0, 0
        o.sn, o.sy, o.sx, o.sc = h * w * c, w * c, c, 1
        o.N, o.H, o.W, o.C = n, h, w, c
        o.bias = bias.data_ptr() if bias is not None else 0
        o.bias_n = bias.numel() if bias is not None else 0
        return o

    def conv_fwd_launch(self, layer, x, ospec, stats=None):
        """stats: fp64 [n, c, 2] tensor the epilogue accumulates (sum, sum of squares) into; returns False when this
        plan's epilogue cannot (the caller then runs the separate statistics kernel)"""
        plan = N.IgemmPlan()
        xs = x.struct()
        L = N.lib()
        N.check(L.aclgan_plan_conv_fwd(C.byref(layer.desc), C.byref(xs), layer.wptr(0), C.byref(ospec),
                                       C.byref(plan)), "plan_conv_fwd")
        fused = stats is not None and self.fuse_stats and bool(L.aclgan_igemm_stats_supported(C.byref(plan)))
        if fused:
            plan.out.stats = stats.data_ptr()
        N.check(L.aclgan_igemm_launch(C.byref(plan), _sp()), "igemm_launch(fwd)")
        return fused

    def conv_out_hw(self, layer, x):
        hp, wp = x.h + 2 * x.pad, x.w + 2 * x.pad
        return (hp - layer.k) // layer.stride + 1, (wp - layer.k) // layer.stride + 1

    def conv_dgrad(self, layer, dy, x):
        """gradient w.r.t. the padded input plane of `layer`: dense [n, hp, wp, cs] in the gradient dtype."""
        hp, wp = x.h + 2 * x.pad, x.w + 2 * x.pad
        cs = x.c if x.c >= 64 else 16           # image planes: 3|6 channels -> one 16-wide UMMA column block
        if x.c < 64:
            g = zeros((x.n, hp, wp, cs), torch.float32, self.device)
        else:
            g = self.new_dense(x.n, hp, wp, cs, zero=(cs != round_up(layer.cin, 16)))
        dys = dy.struct()
        s = layer.stride
        o = self._out_dense(g)
        if s == 2:
            # the four output-parity phases of the transposed stride-2 conv run as ONE launch (phase = -1):
            # `o` describes the phase-(0,0) sub-grid, the plan adds the per-phase offsets
            o.sy, o.sx = 2 * wp * cs, 2 * cs
            o.H, o.W = hp // 2, wp // 2
        plan = N.IgemmPlan()
        N.check(N.lib().aclgan_plan_conv_dgrad(C.byref(layer.desc), C.byref(dys), layer.wptr(1), -1 if s == 2 else 0,
                                               C.byref(o), C.byref(plan)), "plan_conv_dgrad")
        N.check(N.lib().aclgan_igemm_launch(C.byref(plan), _sp()), "igemm_launch(dgrad)")
        return g

    def conv_wgrad(self, layer, dy, x, transpose_taps=False):
        plan = N.WgradPlan()
        dys, xs = dy.struct(), x.struct()
        # transposed taps need a box-per-tap plan (below); the plan builder would pick a segment plan whenever the strip is a
        # multiple of 64 pixels long (2H - 4 = 64 k: inputs of 68, 136, 260 ... pixels), so that choice is switched off for this
        # one host-side call (the switch is read with getenv at plan time; plans are built on one thread, at capture time)
        with _plan_env("ACLGAN_WGRAD_SEG", "0" if transpose_taps else None):
            N.check(N.lib().aclgan_plan_conv_wgrad(C.byref(layer.desc), C.byref(dys), C.byref(xs),
                                                   layer.dw().data_ptr(), C.byref(plan)), "plan_conv_wgrad")
        if transpose_taps:
            # the operands are TRANSPOSED strips convolved with the transposed filter: tap (kh, kw) of the plan is tap (kw, kh)
            # of the stored weight gradient (box-per-tap plans only: the segment plans write consecutive tap slots)
            k = layer.k
            assert plan.seg_mode == 0 and plan.num_taps == k * k
            for t in range(k * k):
                plan.tap_out[t] = (t % k) * k + t // k
        N.check(N.lib().aclgan_wgrad_launch(C.byref(plan), _sp()), "wgrad_launch")

    def dy_pad(self, layer):
        if layer.window == N.WINDOW_OUT or layer.stride == 1:
            return layer.k - 1
        return layer.k // 2 - 1

    # statistics -> coefficients -> element-wise pass.  Two launches by default; ACLGAN_FUSE_FINALIZE=1 folds the coefficient step
    # into the row kernels (every CTA derives its own channels' coefficients: ~190 fewer launches per step-pair, but measured
    # 0.5 ms SLOWER on the 256x256 batch-8 step - the fp64 chain then sits in front of every CTA's first load)
    def _finalize_apply(self, f, a):
        L = N.lib()
        if self.fuse_finalize:
            N.check(L.aclgan_norm_finalize_apply(C.byref(f), C.byref(a), _sp()), "norm_finalize_apply")
        else:
            N.check(L.aclgan_norm_finalize(C.byref(f), _sp()), "norm_finalize")
            N.check(L.aclgan_norm_apply(C.byref(a), _sp()), "norm_apply")

    def _bwd_finalize_apply(self, f, b):
        L = N.lib()
        if self.fuse_finalize:
            return L.aclgan_norm_bwd_finalize_apply(C.byref(f), C.byref(b), _sp())
        N.check(L.aclgan_norm_bwd_finalize(C.byref(f), _sp()), "norm_bwd_finalize")
        return L.aclgan_block_bwd_apply(C.byref(b), _sp())

    # ------------------------------------------------------------------------------------------ conv block
    def conv_block(self, tape, layer, x, norm=N.NORM_NONE, act=N.ACT_NONE, out_pad=0, upsample=1, res=None,
                   adain=None, ln=None, train_w=True):
        """pad -> conv -> norm -> activation (-> + residual) -> [2x nearest upsample] -> reflect pad of the consumer.
        adain = (weight, bias, d_weight, d_bias): fp32 [n, c] views with a common row stride (slices of the MLP output
        row and of its gradient; d_* may be None) ; ln = (gamma, beta, dgamma view, dbeta view)."""
        L = N.lib()
        ho, wo = self.conv_out_hw(layer, x)
        n, cout = x.n, layer.cout
        slope = 0.2
        if norm == N.NORM_NONE:
            assert upsample == 1 and res is None
            out = ActT(self, n, ho, wo, cout, out_pad)
            self.conv_fwd_launch(layer, x, self._out_plane(out, act, layer.bias, slope))
            saved = None
        else:
            cs = round_up(cout, 64)
            y = self.new_dense(n, ho, wo, cs, zero=(cs != cout))
            sums = self.sums(n, cs)
            y4 = self.t4(y)
            if not self.conv_fwd_launch(layer, x, self._out_dense(y, layer.bias), stats=sums):
                N.check(L.aclgan_norm_stats(C.byref(y4), sums.data_ptr(), _sp()), "norm_stats")
            coef = torch.empty((4, n, cs), dtype=torch.float32, device=self.device)   # scale, shift, mean, inv
            sigma = torch.empty((n,), dtype=torch.float32, device=self.device)
            f = N.NormFinalizeArgs()
            f.mode, f.n, f.c, f.hw, f.c_valid, f.eps = norm, n, cs, ho * wo, cout, self.eps
            f.sums = sums.data_ptr()
            if norm == N.NORM_ADAIN:
                assert adain[0].stride(1) == 1 and adain[1].stride() == adain[0].stride()
                f.w, f.b, f.wb_stride = adain[0].data_ptr(), adain[1].data_ptr(), adain[0].stride(0)
            elif norm == N.NORM_LN:
                f.w, f.b = ln[0].data_ptr(), ln[1].data_ptr()
            f.scale, f.shift, f.mean, f.inv = (coef[i].data_ptr() for i in range(4))
            f.sigma = sigma.data_ptr()
            out = ActT(self, n, ho * upsample, wo * upsample, cout, out_pad)
            a = N.ApplyArgs()
            a.y, a.scale, a.shift, a.act, a.slope = y4, coef[0].data_ptr(), coef[1].data_ptr(), act, slope
            a.has_res = 1 if res is not None else 0
            if res is not None:
                a.res = res.struct()
            a.upsample, a.dst = upsample, out.struct()
            self._finalize_apply(f, a)
            saved = (y, coef, sigma, sums if norm == N.NORM_LN else None)     # (pool slices live until the update ends)
        need_x_grad = x.requires_grad
        out.requires_grad = need_x_grad or train_w or (res is not None and res.requires_grad) or adain is not None
        if not tape.enabled or not out.requires_grad:
            return out

        def bwd():
            gp, gr = out.gp, out.gr
            out.gp = out.gr = None
            if gp is None and gr is None:
                return
            if res is not None and res.requires_grad:
                # gradient of the residual branch = gradient of the block output (dense, un-padded)
                res.add_gr(self._fold_only(gp, gr, out, upsample))
            b = N.BlockBwdArgs()
            b.gp = gp.data_ptr() if gp is not None else 0
            b.gr = gr.data_ptr() if gr is not None else 0
            b.g_kind, b.gp_pad, b.upsample = self.prec.kind, out.pad, upsample
            cs = round_up(cout, 64)
            b.n, b.h, b.w, b.c = n, ho, wo, cs
            b.slope = 0.0 if act == N.ACT_RELU else slope
            dy = ActT(self, n, ho, wo, cout, self.dy_pad(layer))
            b.dy = dy.struct()
            if norm == N.NORM_NONE:
                b.norm = 0
                b.mask_mode = N.MASK_NONE if act == N.ACT_NONE else N.MASK_FROM_OUT
                b.out = out.struct()
                if train_w:         # conv bias gradient = sum of dz: fused into the apply pass below
                    b.dbias, b.dbias_n = layer.db().data_ptr(), cout
            else:
                y, coef, sigma, fwd_sums = saved
                sums = self.sums(n, cs)
                b.sums = sums.data_ptr()
                b.norm = 1
                b.mask_mode = N.MASK_NONE if act == N.ACT_NONE else N.MASK_FROM_Z
                b.y = self.t4(y)
                b.scale, b.shift, b.mean, b.inv = (coef[i].data_ptr() for i in range(4))
                N.check(L.aclgan_block_bwd_reduce(C.byref(b), _sp()), "block_bwd_reduce")
                cf = torch.empty((3, n, cs), dtype=torch.float32, device=self.device)
                f = N.NormBwdFinalizeArgs()
                f.mode, f.n, f.c, f.hw, f.c_valid = norm, n, cs, ho * wo, cout
                f.sums, f.inv, f.sigma = sums.data_ptr(), coef[3].data_ptr(), sigma.data_ptr()
                if norm == N.NORM_ADAIN:
                    assert adain[2].stride() == adain[0].stride() and adain[3].stride() == adain[0].stride()
                    f.w, f.dw, f.db = adain[0].data_ptr(), adain[2].data_ptr(), adain[3].data_ptr()
                    f.wb_stride = adain[0].stride(0)
                elif norm == N.NORM_LN:
                    f.w, f.dw, f.db = ln[0].data_ptr(), ln[2].data_ptr(), ln[3].data_ptr()
                f.ca, f.cb, f.cc = (cf[i].data_ptr() for i in range(3))
                if norm == N.NORM_LN and train_w:
                    # a conv bias in front of LayerNorm is NOT cancelled (statistics span all channels): the finalize kernel
                    # adds db[c] = sum_n ca*T1 + cb*sum_hw(yhat) + cc*HW with sum_hw(yhat) = (S1 - HW*mean) * inv
                    f.fsums, f.mean, f.dbias = fwd_sums.data_ptr(), coef[2].data_ptr(), layer.db().data_ptr()
                b.ca, b.cb, b.cc = (cf[i].data_ptr() for i in range(3))
            if norm == N.NORM_NONE:
                rc = L.aclgan_block_bwd_apply(C.byref(b), _sp())
            else:
                rc = self._bwd_finalize_apply(f, b)
            if rc == -3 and b.dbias:
                # degenerate plane sizes (generic kernels): separate reduction for the bias gradient
                b.dbias = 0
                sums = self.sums(n, cs)
                b.sums = sums.data_ptr()
                N.check(L.aclgan_block_bwd_reduce(C.byref(b), _sp()), "block_bwd_reduce")
                N.check(L.aclgan_stats_to_bias(sums.data_ptr(), layer.db().data_ptr(), n, cs, cout, _sp()), "stats_to_bias")
                rc = L.aclgan_block_bwd_apply(C.byref(b), _sp())
            N.check(rc, "block_bwd_apply")
            if self.debug is not None:
                self.debug.setdefault(id(layer), []).append(dict(
                    dy=dy.value_nchw(), gp=None if gp is None else gp.float().clone(),
                    gr=None if gr is None else gr.float().clone(), out_pad=out.pad, upsample=upsample,
                    xpad=x.buf[:, :x.numel].float().sum(0).view(x.n, x.h + 2 * x.pad, x.w + 2 * x.pad, x.c).clone()))
            if train_w:
                self.conv_wgrad(layer, dy, x)
            if need_x_grad:
                x.add_gp(self.conv_dgrad(layer, dy, x))

        tape.push(bwd)
        return out

    # ------------------------------------------------------------------------------------------ sub-pixel up block
    def conv_block_up(self, tape, up, x, act=N.ACT_RELU, out_pad=0, ln=None, train_w=True):
        """nearest 2x upsample -> reflect pad 2 -> conv 5x5 -> LayerNorm -> activation -> reflect pad of the consumer
        (reference networks.py:256-257 + :520-536) with the up-sampled plane never materialised: `x` is the SOURCE plane
        (reflect pad 1).  Main 3x3 / 4*Cout convolution + depth-to-space epilogue, exact 5x5 ring strips, LayerNorm
        statistics fused into the three conv epilogues (csrc/upconv.cu header; tools/subpixel_pipeline.py)."""
        L = N.lib()
        base = up.base
        assert x.pad == 1 and x.h >= 3 and x.w >= 3
        n, H, W, cout = x.n, x.h, x.w, base.cout
        cs = round_up(cout, 64)
        row = 2 * W * cs                                     # elements per output row of y
        y = self.new_dense(n, 2 * H, 2 * W, cs, zero=(cs != cout))
        slope = 0.2

        def spec(N_, H_, W_, C_, sy, sx, off=0, bias=None, bias_n=0):
            o = N.OutSpec()
            o.ptr[0] = y.data_ptr()
            o.kind = N.OUT_BF16 if y.dtype == torch.bfloat16 else N.OUT_F32
            o.act, o.slope, o.mirror, o.off = N.ACT_NONE, 0.2, 0, off
            o.sn, o.sy, o.sx, o.sc = 2 * H * row, sy, sx, 1
            o.N, o.H, o.W, o.C = N_, H_, W_, C_
            o.bias, o.bias_n = (bias.data_ptr() if bias is not None else 0), bias_n
            return o

        # ---- main: 3x3 conv of the source plane, 4 phases folded into N, depth-to-space store, ring pixels skipped
        om = spec(n, H, W, 4 * cout, 2 * row, 2 * cs, bias=up.bias4, bias_n=4 * cout)
        om.d2s_c, om.d2s_sy, om.d2s_sx, om.ring = cout, row, cs, 1
        plan = N.IgemmPlan()
        xs = x.struct()
        N.check(L.aclgan_plan_conv_fwd(C.byref(up.desc), C.byref(xs), up.wptr(0), C.byref(om), C.byref(plan)), "plan_conv_fwd(up)")
        fused = self.fuse_stats and bool(L.aclgan_igemm_stats_supported(C.byref(plan)))
        groups = 4 if fused else 1
        sums = self.sums(n, 4 * cout if fused else cs)
        if fused:
            plan.out.stats = sums.data_ptr()
        N.check(L.aclgan_igemm_launch(C.byref(plan), _sp()), "igemm_launch(up main)")
        # ---- ring: exact 5x5 on strips of the padded up-sampled plane (columns as transposed strips / transposed filter)
        xr = ActT(self, 2 * n, 2, 2 * W, x.c_valid, 2, cs=x.c)
        xc = ActT(self, 2 * n, 2, 2 * H - 4, x.c_valid, 2, cs=x.c)
        g = N.UpStripsArgs()
        g.src, g.rows, g.cols = xs, xr.struct(), xc.struct()
        N.check(L.aclgan_up_gather_strips(C.byref(g), _sp()), "up_gather_strips")
        orow = spec(2 * n, 2, 2 * W, cs, row, cs, bias=base.bias, bias_n=cout)
        orow.z_mod, orow.z_off = n, (2 * H - 2) * row
        ocol = spec(2 * n, 2, 2 * H - 4, cs, cs, row, off=2 * row, bias=base.bias, bias_n=cout)
        ocol.z_mod, ocol.z_off = n, (2 * W - 2) * cs
        for lay, xin, o in ((base, xr, orow), (up.colview, xc, ocol)):
            if fused:
                o.stats_c = 4 * cout
            if not self.conv_fwd_launch(lay, xin, o, stats=sums if fused else None) and fused:
                raise N.NativeError("up-block strip convolution cannot fuse the LayerNorm statistics")
        y4 = self.t4(y)
        if not fused:
            N.check(L.aclgan_norm_stats(C.byref(y4), sums.data_ptr(), _sp()), "norm_stats")
        # ---- LayerNorm + activation -> the consumer's padded plane
        coef = torch.empty((4, n, cs), dtype=torch.float32, device=self.device)
        sigma = torch.empty((n,), dtype=torch.float32, device=self.device)
        f = N.NormFinalizeArgs()
        f.mode, f.n, f.c, f.hw, f.c_valid, f.eps = N.NORM_LN, n, cs, 4 * H * W, cout, self.eps
        f.sums, f.stat_groups = sums.data_ptr(), groups
        f.w, f.b = ln[0].data_ptr(), ln[1].data_ptr()
        f.scale, f.shift, f.mean, f.inv = (coef[i].data_ptr() for i in range(4))
        f.sigma = sigma.data_ptr()
        out = ActT(self, n, 2 * H, 2 * W, cout, out_pad)
        a = N.ApplyArgs()
        a.y, a.scale, a.shift, a.act, a.slope = y4, coef[0].data_ptr(), coef[1].data_ptr(), act, slope
        a.has_res, a.upsample, a.dst = 0, 1, out.struct()
        self._finalize_apply(f, a)
        need_x_grad = x.requires_grad
        out.requires_grad = need_x_grad or train_w
        if not tape.enabled or not out.requires_grad:
            return out

        def bwd():
            gp, gr = out.gp, out.gr
            out.gp = out.gr = None
            if gp is None and gr is None:
                return
            b = N.BlockBwdArgs()
            b.gp = gp.data_ptr() if gp is not None else 0
            b.gr = gr.data_ptr() if gr is not None else 0
            b.g_kind, b.gp_pad, b.upsample = self.prec.kind, out.pad, 1
            b.n, b.h, b.w, b.c = n, 2 * H, 2 * W, cs
            b.slope = 0.0 if act == N.ACT_RELU else slope
            dy = ActT(self, n, 2 * H, 2 * W, cout, 0)
            b.dy = dy.struct()
            bs = self.sums(n, cs)
            b.sums, b.norm = bs.data_ptr(), 1
            b.mask_mode = N.MASK_NONE if act == N.ACT_NONE else N.MASK_FROM_Z
            b.y = self.t4(y)            # (referencing `y` here keeps the raw conv output alive until the backward pass)
            b.scale, b.shift, b.mean, b.inv = (coef[i].data_ptr() for i in range(4))
            N.check(L.aclgan_block_bwd_reduce(C.byref(b), _sp()), "block_bwd_reduce")
            cf = torch.empty((3, n, cs), dtype=torch.float32, device=self.device)
            fb = N.NormBwdFinalizeArgs()
            fb.mode, fb.n, fb.c, fb.hw, fb.c_valid = N.NORM_LN, n, cs, 4 * H * W, cout
            fb.sums, fb.inv, fb.sigma = bs.data_ptr(), coef[3].data_ptr(), sigma.data_ptr()
            fb.w, fb.dw, fb.db = ln[0].data_ptr(), ln[2].data_ptr(), ln[3].data_ptr()
            fb.ca, fb.cb, fb.cc = (cf[i].data_ptr() for i in range(3))
            if train_w:         # conv bias in front of the LayerNorm (not cancelled): from the forward statistics
                fb.fsums, fb.mean, fb.dbias, fb.fstat_groups = sums.data_ptr(), coef[2].data_ptr(), base.db().data_ptr(), groups
            b.ca, b.cb, b.cc = (cf[i].data_ptr() for i in range(3))
            N.check(self._bwd_finalize_apply(fb, b), "block_bwd_apply")
            # dY -> space-to-depth plane (ring zeroed) for the 3x3 main gradients + ring strips for the 5x5 strip gradients
            ds = ActT(self, n, H, W, 4 * cout, 2)
            dr = ActT(self, 2 * n, 2, 2 * W, cout, 4)
            dc = ActT(self, 2 * n, 2, 2 * H - 4, cout, 4)
            pk = N.UpDyPackArgs()
            pk.dy, pk.cout, pk.s2d, pk.rows, pk.cols = dy.struct(), cout, ds.struct(), dr.struct(), dc.struct()
            N.check(L.aclgan_up_dy_pack(C.byref(pk), _sp()), "up_dy_pack")
            if train_w:
                zero_(up.dwp)
                self.conv_wgrad(up.main, ds, x)
                fw = N.UpFoldWgradArgs()
                fw.dwp, fw.dw5, fw.co, fw.ci = up.dwp.data_ptr(), base.arena.flat.data_ptr(), cout, base.cin
                for j in range(5):
                    fw.affp[j] = up.aff[up.layout][j]
                    fw.aff5[j] = base.aff[base.layout][j] + (base.dw_off if j == 0 else 0)
                N.check(L.aclgan_up_fold_wgrad(C.byref(fw), _sp()), "up_fold_wgrad")
                self.conv_wgrad(base, dr, xr)
                self.conv_wgrad(up.colview, dc, xc, transpose_taps=True)
            if need_x_grad:
                gx = self.conv_dgrad(up.main, ds, x)
                gxr = self.conv_dgrad(base, dr, xr)
                gxc = self.conv_dgrad(up.colview, dc, xc)
                sc = N.UpScatterArgs()
                sc.g, sc.grows, sc.gcols = gx.data_ptr(), gxr.data_ptr(), gxc.data_ptr()
                sc.kind, sc.n, sc.h, sc.w, sc.c = self.prec.kind, n, H, W, gx.shape[-1]
                assert gxr.shape[-1] == gx.shape[-1] and gxc.shape[-1] == gx.shape[-1]
                N.check(L.aclgan_up_scatter_strips(C.byref(sc), _sp()), "up_scatter_strips")
                x.add_gp(gx)

        tape.push(bwd)
        return out

    def _fold_only(self, gp, gr, out, upsample):
        """dense gradient of the logical block output (fold of the padded / upsampled plane gradient)"""
        if gp is None:
            return copy_(torch.empty_like(gr), gr)
        b = N.BlockBwdArgs()
        cs = gp.shape[-1]
        h, w = out.h // upsample, out.w // upsample
        b.gp, b.gr = gp.data_ptr(), (gr.data_ptr() if gr is not None else 0)
        b.g_kind, b.gp_pad, b.upsample = self.prec.kind, out.pad, upsample
        b.n, b.h, b.w, b.c = out.n, h, w, cs
        b.mask_mode, b.norm = N.MASK_NONE, 0
        # block_bwd_apply writes bf16 plane(s); a dense gradient in the engine dtype is wanted here, so the fold
        # goes through the plane writer and is read back as hi(+lo)
        tmp = ActT(self, out.n, h, w, out.c_valid, 0)
        b.dy = tmp.struct()
        N.check(N.lib().aclgan_block_bwd_apply(C.byref(b), _sp()), "block_bwd_apply(fold)")
        v = tmp.buf[:, :tmp.numel].float().sum(0) if self.prec.planes == 2 else tmp.buf[0, :tmp.numel]
        return v.view(out.n, h, w, cs).to(self.prec.dtype)

    # ------------------------------------------------------------------------------------------ final conv
    def conv_to_image(self, tape, layer, x, act=N.ACT_TANH, train_w=True):
        """last decoder block: conv -> tanh -> NCHW fp32 image (networks.py:260)"""
        ho, wo = self.conv_out_hw(layer, x)
        n, cout = x.n, layer.cout
        img = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=self.device)
        o = N.OutSpec()
        o.ptr[0] = img.data_ptr()
        o.kind, o.act, o.slope, o.mirror, o.off = N.OUT_F32, act, 0.2, 0, 0
        o.sn, o.sy, o.sx, o.sc = cout * ho * wo, wo, 1, ho * wo
        o.N, o.H, o.W, o.C = n, ho, wo, cout
        o.bias = layer.bias.data_ptr()
        o.bias_n = cout
        self.conv_fwd_launch(layer, x, o)
        out = ImgT(img, requires_grad=tape.enabled)
        if not tape.enabled:
            return out

        def bwd():
            if out.grad is None:
                return
            if out.grad.is_cuda:        # assembled on the caller's stream (trainer._slice), consumed on this chain's
                out.grad.record_stream(torch.cuda.current_stream())
            dy = ActT(self, n, ho, wo, cout, layer.k - 1, cs=8, zero=True)
            a = N.ImgGradPackArgs()
            a.dimg = out.grad.data_ptr()
            a.out_img = img.data_ptr() if act == N.ACT_TANH else 0
            a.n, a.c, a.h, a.w = n, cout, ho, wo
            a.dy = dy.struct()
            a.dbias = layer.db().data_ptr() if train_w else 0
            N.check(N.lib().aclgan_img_grad_pack(C.byref(a), _sp()), "img_grad_pack")
            out.grad = None
            if train_w:
                self.conv_wgrad(layer, dy, x)
            if x.requires_grad:
                x.add_gp(self.conv_dgrad(layer, dy, x))

        tape.push(bwd)
        return out
