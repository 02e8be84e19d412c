"""CPU emulation of the device-side contract of the implicit-GEMM plans (test infrastructure).

`aclgan_plan_*` (host C++) turns a convolution into tensor-map specs + filter-tap box offsets.  The CUDA
kernels execute such a plan with TMA box loads (zero fill out of bounds) and tcgen05 MMAs.  This module
executes the SAME plan on CPU tensors - box loads become index gathers, the MMA a float64 matmul, the
epilogue the documented store rule - so the geometry is validated without a GPU.
"""
import ctypes as C

import torch

import aclgan_native as N


class Memory:
    """Maps 'device addresses' (here: host data_ptr()s of CPU tensors) back to tensors."""

    def __init__(self):
        self.bufs = []

    def add(self, t):
        assert t.is_contiguous()
        self.bufs.append((t.data_ptr(), t.numel() * t.element_size(), t))
        return t.data_ptr()

    def find(self, addr):
        for base, nbytes, t in self.bufs:
            if base <= addr < base + nbytes:
                return t, (addr - base) // t.element_size()
        raise KeyError("address %x not registered" % addr)


def tma_box(mem, spec, coords):
    """Returns the box [.., box1, box0] as float64 (zeros where any coordinate is out of bounds)."""
    rank = spec.rank
    t, off0 = mem.find(spec.base)
    flat = t.reshape(-1)
    es = spec.elem_bytes
    idx = torch.zeros([spec.box[i] for i in reversed(range(rank))], dtype=torch.int64)
    ok = torch.ones_like(idx, dtype=torch.bool)
    for i in range(rank):
        ar = torch.arange(spec.box[i], dtype=torch.int64) + int(coords[i])
        shape = [1] * rank
        shape[rank - 1 - i] = spec.box[i]
        ar = ar.view(shape)
        ok = ok & (ar >= 0) & (ar < int(spec.dims[i]))
        assert spec.strides[i] % es == 0
        idx = idx + ar * (int(spec.strides[i]) // es)
    idx = idx + off0
    safe = idx.clamp(0, flat.numel() - 1)
    assert bool(((idx == safe) | ~ok).all()), "in-bounds TMA coordinate maps outside the allocation"
    vals = flat[safe].double()
    return torch.where(ok, vals, torch.zeros_like(vals))


def _mirror(c, L, p):
    out = [c]
    if p > 0:
        if 1 <= c <= p:
            out.append(-c)
        if L - 1 - p <= c <= L - 2:
            out.append(2 * (L - 1) - c)
    return out


def _act(v, act, slope):
    if act == N.ACT_RELU:
        return torch.relu(v)
    if act == N.ACT_LRELU:
        return torch.where(v > 0, v, v * slope)
    if act == N.ACT_TANH:
        return torch.tanh(v)
    return v


def run_igemm(mem, plan, use_seg=None):
    """Executes an IgemmPlan on CPU buffers registered in `mem` (writes the output buffers in place).
    use_seg: execute the segment description of the plan (what igemm_seg_kernel does: one staged block of seg_rows
    pixels per filter row and chunk, taps = row-shifted 128-row windows of it) instead of the box-per-tap one;
    default: whenever the plan has one."""
    if use_seg is None:
        use_seg = bool(plan.seg_mode)
    assert not use_seg or (plan.seg_mode and plan.num_segs * plan.seg_taps == plan.num_taps)
    o = plan.out
    outs = []
    nplanes_out = 2 if o.kind == N.OUT_SPLIT else 1
    for pl in range(nplanes_out):
        outs.append(mem.find(o.ptr[pl]))
    bias = None
    if o.bias:
        bt, boff = mem.find(o.bias)
        bias = bt.reshape(-1)[boff:].double()
    segs = [(0, 0)] if plan.nseg == 1 else [(0, 0), (0, 1), (1, 0)]
    bn = plan.block_n
    if plan.fold:
        return _run_igemm_fold(mem, plan, outs, bias, segs)
    n_groups = plan.n_groups if plan.n_groups > 1 else 1
    gtaps = plan.group_taps if plan.n_groups > 1 else plan.num_taps
    for grp in range(n_groups):
      for tz in range(plan.tiles_z):
        for ty in range(plan.tiles_y):
            for tx in range(plan.tiles_x):
                x0, y0, z0 = tx * plan.box_x, ty * plan.box_y, tz * plan.box_z
                for nt in range(plan.n_tiles):
                    acc = torch.zeros(128, bn, dtype=torch.float64)
                    for pa, pb in segs:
                        if use_seg:
                            for cc in range(plan.cchunks):
                                for sg in range(plan.num_segs):
                                    S = tma_box(mem, plan.a_seg[pa], (cc * 64, x0 + plan.seg_dx[sg], y0 + plan.seg_dy[sg], z0))
                                    S = S.reshape(plan.seg_rows, 64)
                                    for j in range(plan.seg_taps):
                                        t = sg * plan.seg_taps + j
                                        A = S[plan.tap_row[t]:plan.tap_row[t] + 128]
                                        B = tma_box(mem, plan.b[pb], (plan.tap_bk[t] + cc * 64, nt * bn)).reshape(bn, 64)
                                        acc += A @ B.t()
                            continue
                        for t in range(grp * gtaps, (grp + 1) * gtaps):
                            am = plan.a[pa][plan.tap_var[t]]
                            for cc in range(plan.cchunks):
                                A = tma_box(mem, am, (cc * 64, x0 + plan.tap_dx[t], y0 + plan.tap_dy[t], z0))
                                A = A.reshape(128, 64)
                                B = tma_box(mem, plan.b[pb], (plan.tap_bk[t] + cc * 64, nt * bn)).reshape(bn, 64)
                                acc += A @ B.t()
                    for r in range(128):
                        if plan.flat:
                            q = tx * plan.box_x + r
                            z, rem = divmod(q, plan.flat_img)
                            y, x = divmod(rem, plan.flat_w)
                        else:
                            x = x0 + r % plan.box_x
                            y = y0 + (r // plan.box_x) % plan.box_y
                            z = z0 + r // (plan.box_x * plan.box_y)
                        if not (x < o.W and y < o.H and z < o.N):
                            continue
                        ch0 = nt * bn
                        cnt = min(bn, o.C - ch0)
                        if cnt <= 0:
                            continue
                        v = acc[r, :cnt].clone()
                        if bias is not None:
                            nb = max(0, min(cnt, o.bias_n - ch0))
                            v[:nb] = v[:nb] + bias[ch0:ch0 + nb]
                        v = _act(v, o.act, o.slope)
                        for yy in _mirror(y, o.H, o.mirror):
                            for xx in _mirror(x, o.W, o.mirror):
                                pix = o.off + (plan.group_off[grp] if n_groups > 1 else 0) + z * o.sn + yy * o.sy + xx * o.sx
                                for pl in range(nplanes_out):
                                    t, toff = outs[pl]
                                    flat = t.reshape(-1)
                                    idx = toff + pix + (ch0 + torch.arange(cnt)) * o.sc
                                    if o.kind == N.OUT_SPLIT:
                                        hi = v.float().bfloat16()
                                        val = hi if pl == 0 else (v.float() - hi.float()).bfloat16()
                                    else:
                                        val = v
                                    if o.kind == N.OUT_F32_ATOMIC:
                                        flat[idx] += val.to(flat.dtype)
                                    else:
                                        flat[idx] = val.to(flat.dtype)


def _run_igemm_fold(mem, plan, outs, bias, segs):
    """fold mode: P[r][kw*8 + co] per 128-row tile (tiles step by tile_step), out[q][co] = sum_kw P[q + kw][kw*8 + co]"""
    o = plan.out
    k = plan.fold
    assert plan.flat and plan.block_n == 64 and plan.tile_step + k - 1 <= 128 and o.kind in (N.OUT_F32, N.OUT_BF16)
    for tx in range(plan.tiles_x):
        x0 = tx * plan.tile_step
        acc = torch.zeros(128, 64, dtype=torch.float64)
        for pa, pb in segs:
            for t in range(plan.num_taps):
                for cc in range(plan.cchunks):
                    A = tma_box(mem, plan.a[pa][0], (cc * 64, x0 + plan.tap_dx[t], 0, 0)).reshape(128, 64)
                    B = tma_box(mem, plan.b[pb], (plan.tap_bk[t] + cc * 64, 0)).reshape(64, 64)
                    acc += A @ B.t()
        for r in range(plan.tile_step):
            q = x0 + r
            z, rem = divmod(q, plan.flat_img)
            y, x = divmod(rem, plan.flat_w)
            if not (x < o.W and y < o.H and z < o.N):
                continue
            v = torch.stack([sum(acc[r + kw, kw * 8 + co] for kw in range(k)) for co in range(o.C)])
            if bias is not None:
                nb = min(o.C, o.bias_n)
                v[:nb] = v[:nb] + bias[:nb]
            v = _act(v, o.act, o.slope)
            t_, toff = outs[0]
            flat = t_.reshape(-1)
            idx = toff + o.off + z * o.sn + y * o.sy + x * o.sx + torch.arange(o.C) * o.sc
            flat[idx] = v.to(flat.dtype)


def pack_weight(desc, w_oihw, transposed, dtype=torch.bfloat16):
    """Packs an OIHW weight into the layout the plans expect.  The C index function is affine in
    (co, ci, kh, kw); its coefficients are probed and the scatter is vectorised (checked on random samples)."""
    L = N.lib()
    rows, kt = C.c_int64(), C.c_int64()
    L.aclgan_packed_weight_shape(C.byref(desc), int(transposed), C.byref(rows), C.byref(kt))
    co, ci, kh, kw = w_oihw.shape

    def idx(a, b, c, d):
        return L.aclgan_packed_weight_index(C.byref(desc), int(transposed), a, b, c, d)

    base = idx(0, 0, 0, 0)
    s_co = idx(1, 0, 0, 0) - base if co > 1 else 0
    s_ci = idx(0, 1, 0, 0) - base if ci > 1 else 0
    s_kh = idx(0, 0, 1, 0) - base if kh > 1 else 0
    s_kw = idx(0, 0, 0, 1) - base if kw > 1 else 0
    g = torch.Generator().manual_seed(0)
    for _ in range(16):
        a, b, c, d = [int(torch.randint(0, n, (1,), generator=g)) for n in (co, ci, kh, kw)]
        assert idx(a, b, c, d) == base + a * s_co + b * s_ci + c * s_kh + d * s_kw
    ii = (base + torch.arange(co).view(-1, 1, 1, 1) * s_co + torch.arange(ci).view(1, -1, 1, 1) * s_ci +
          torch.arange(kh).view(1, 1, -1, 1) * s_kh + torch.arange(kw).view(1, 1, 1, -1) * s_kw)
    out = torch.zeros(rows.value * kt.value, dtype=torch.float32)
    out[ii.reshape(-1)] = w_oihw.reshape(-1).float()
    return out.reshape(rows.value, kt.value).to(dtype)


def run_wgrad(mem, plan):
    """Executes a WgradPlan on CPU buffers: dw[m][tap][n] += sum_pixels Mop[pixel][m] * Nop[pixel][n]."""
    dwt, dwoff = mem.find(plan.dw)
    dw = dwt.reshape(-1)
    segs = [(0, 0)] if plan.nseg == 1 else [(0, 0), (0, 1), (1, 0)]
    ncols = 64 * plan.n_chunks
    pix = plan.box_x * plan.box_y * plan.box_z
    if plan.seg_mode:
        return _run_wgrad_seg(mem, plan, dw, dwoff, segs, ncols)
    for tap in range(plan.num_taps):
        for mt in range(plan.m_tiles):
            for nt in range(plan.n_tiles):
                acc = torch.zeros(128, ncols, dtype=torch.float64)
                for bz in range(plan.blocks_z):
                    for by in range(plan.blocks_y):
                        for bx in range(plan.blocks_x):
                            x0, y0, z0 = bx * plan.box_x, by * plan.box_y, bz * plan.box_z
                            for pm, pn in segs:
                                mm = plan.mop[pm][plan.m_var[tap]]
                                nm = plan.nop[pn][plan.n_var[tap]]
                                Mt = torch.zeros(pix, 128, dtype=torch.float64)
                                for c in range(plan.m_chunks):
                                    Mt[:, c * 64:(c + 1) * 64] = tma_box(
                                        mem, mm, ((mt * 2 + c) * 64, x0 + plan.m_dx[tap], y0 + plan.m_dy[tap], z0)).reshape(pix, 64)
                                Nt = torch.zeros(pix, ncols, dtype=torch.float64)
                                for c in range(plan.n_chunks):
                                    Nt[:, c * 64:(c + 1) * 64] = tma_box(
                                        mem, nm, ((nt * plan.n_chunks + c) * 64, x0 + plan.n_dx[tap], y0 + plan.n_dy[tap], z0)).reshape(pix, 64)
                                acc += Mt.t() @ Nt
                for r in range(128):
                    m = mt * 128 + r
                    if m >= plan.M:
                        continue
                    n0 = nt * ncols
                    cnt = min(ncols, plan.Nn - n0)
                    if cnt <= 0:
                        continue
                    idx = dwoff + m * plan.dw_sm + plan.tap_out[tap] * plan.dw_st + n0 + torch.arange(cnt)
                    dw[idx] += acc[r, :cnt].to(dw.dtype)


def _run_wgrad_seg(mem, plan, dw, dwoff, segs, ncols):
    """segment mode (wgrad_seg_kernel): per filter row one CTA-tile; the shifted operand is staged once per 64-pixel
    block as seg_rows pixels and tap kw uses its rows [kw, kw + 64)"""
    assert plan.box_x * plan.box_y == 64 and plan.box_z == 1
    step = plan.seg_step if plan.seg_step > 0 else 1
    for tap in range(plan.num_taps):
        for mt in range(plan.m_tiles):
            for nt in range(plan.n_tiles):
                kw0, cnt = plan.seg_kw0[tap], plan.seg_cnt[tap]
                assert 1 <= cnt <= plan.seg_taps and cnt * ncols <= 512
                acc = torch.zeros(cnt, 128, ncols, dtype=torch.float64)
                for bz in range(plan.blocks_z):
                    for by in range(plan.blocks_y):
                        for bx in range(plan.blocks_x):
                            x0, y0, z0 = bx * plan.box_x, by * plan.box_y, bz
                            for pm, pn in segs:
                                mm = plan.seg_map[pm] if plan.seg_on_m else plan.mop[pm][0]
                                nm = plan.nop[pn][0] if plan.seg_on_m else plan.seg_map[pn]
                                mr = plan.seg_rows if plan.seg_on_m else 64
                                nr = 64 if plan.seg_on_m else plan.seg_rows
                                Mt = torch.zeros(mr, 128, dtype=torch.float64)
                                for c in range(plan.m_chunks):
                                    Mt[:, c * 64:(c + 1) * 64] = tma_box(
                                        mem, mm, ((mt * 2 + c) * 64, x0 + plan.m_dx[tap], y0 + plan.m_dy[tap], z0)).reshape(mr, 64)
                                Nt = torch.zeros(nr, ncols, dtype=torch.float64)
                                for c in range(plan.n_chunks):
                                    Nt[:, c * 64:(c + 1) * 64] = tma_box(
                                        mem, nm, ((nt * plan.n_chunks + c) * 64, x0 + plan.n_dx[tap], y0 + plan.n_dy[tap], z0)).reshape(nr, 64)
                                for j in range(cnt):
                                    kw = (kw0 + j) * step
                                    Mw = Mt[kw:kw + 64] if plan.seg_on_m else Mt
                                    Nw = Nt if plan.seg_on_m else Nt[kw:kw + 64]
                                    acc[j] += Mw.t() @ Nw
                for kw in range(cnt):
                    for r in range(128):
                        m = mt * 128 + r
                        if m >= plan.M:
                            continue
                        n0 = nt * ncols
                        cnt = min(ncols, plan.Nn - n0)
                        if cnt <= 0:
                            continue
                        idx = dwoff + m * plan.dw_sm + (plan.tap_out[tap] + kw) * plan.dw_st + n0 + torch.arange(cnt)
                        dw[idx] += acc[kw, r, :cnt].to(dw.dtype)


def unpack_wgrad(desc, dw_flat, shape_oihw):
    """fp32 dW buffer (packed-weight layout given by aclgan_wgrad_layout) -> OIHW tensor."""
    L = N.lib()
    transposed = L.aclgan_wgrad_layout(C.byref(desc))
    co, ci, kh, kw = shape_oihw

    def idx(a, b, c, d):
        return L.aclgan_packed_weight_index(C.byref(desc), transposed, a, b, c, d)

    base = idx(0, 0, 0, 0)
    s_co = idx(1, 0, 0, 0) - base if co > 1 else 0
    s_ci = idx(0, 1, 0, 0) - base if ci > 1 else 0
    s_kh = idx(0, 0, 1, 0) - base if kh > 1 else 0
    s_kw = idx(0, 0, 0, 1) - base if kw > 1 else 0
    ii = (base + torch.arange(co).view(-1, 1, 1, 1) * s_co + torch.arange(ci).view(1, -1, 1, 1) * s_ci +
          torch.arange(kh).view(1, 1, -1, 1) * s_kh + torch.arange(kw).view(1, 1, 1, -1) * s_kw)
    return dw_flat.reshape(-1)[ii.to(dw_flat.device)]
