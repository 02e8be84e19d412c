"""Helpers for the -m gpu tests: build padded NHWC bf16 planes on the device and call the C ABI."""
import ctypes as C

import torch
import torch.nn.functional as F

import aclgan_native as N
import emul


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_act(x_nchw, pad, cs, planes=1, slack=64, mode="reflect", dev="cuda"):
    n, c, h, w = x_nchw.shape
    xp = F.pad(x_nchw, (pad,) * 4, mode=mode) if pad > 0 else x_nchw
    v = torch.zeros(n, h + 2 * pad, w + 2 * pad, cs, device=xp.device)
    v[..., :c] = xp.permute(0, 2, 3, 1)
    buf = torch.zeros(planes, v.numel() + slack, dtype=torch.bfloat16, device=dev)
    hi = v.bfloat16()
    buf[0, :v.numel()] = hi.reshape(-1).to(dev)
    if planes == 2:
        buf[1, :v.numel()] = (v - hi.float()).bfloat16().reshape(-1).to(dev)
    act = N.Act()
    for p in range(planes):
        act.data[p] = buf[p].data_ptr()
    act.planes, act.n, act.h, act.w, act.c, act.pad = planes, n, h, w, cs, pad
    eff = buf[:, :v.numel()].float().sum(0).reshape(n, h + 2 * pad, w + 2 * pad, cs)[..., :c].permute(0, 3, 1, 2)
    return act, buf, eff.double()


def split_planes(t, planes):
    hi = t.bfloat16()
    if planes == 1:
        return [hi], hi.double()
    lo = (t - hi.float()).bfloat16()
    return [hi, lo], hi.double() + lo.double()


def pack_weights(desc, w, transposed, planes, dev="cuda"):
    """returns (list of packed device tensors, effective fp64 OIHW weight the tensor cores see)"""
    wc = w.detach().float().cpu()
    parts, eff = split_planes(wc, planes)
    packed = [emul.pack_weight(desc, p.float(), transposed).to(dev) for p in parts]
    return packed, eff


def wptr(packed):
    return (C.c_uint64 * 2)(*([t.data_ptr() for t in packed] + [0] * (2 - len(packed))))


def out_spec(n, h, w, c, kind=N.OUT_F32, pad=0, act=N.ACT_NONE, bias=None, mirror=0, dev="cuda"):
    hp, wp = h + 2 * pad, w + 2 * pad
    dt = torch.float32 if kind in (N.OUT_F32, N.OUT_F32_ATOMIC) else torch.bfloat16
    planes = 2 if kind == N.OUT_SPLIT else 1
    buf = torch.zeros(planes, n, hp, wp, c, dtype=dt, device=dev)
    o = N.OutSpec()
    for p in range(planes):
        o.ptr[p] = buf[p].data_ptr()
    o.kind, o.act, o.slope, o.mirror = kind, act, 0.2, mirror
    o.off = (pad * wp + pad) * c
    o.sn, o.sy, o.sx, o.sc = hp * wp * c, wp * c, c, 1
    o.N, o.H, o.W, o.C = n, h, w, c
    if bias is not None:
        o.bias = bias.data_ptr()
        o.bias_n = bias.numel()
    return o, buf


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))
