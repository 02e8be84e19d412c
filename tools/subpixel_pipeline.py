"""Pure-torch (fp64) model of the sub-pixel up-convolution pipeline of csrc/upconv.cu + engine.conv_block_up, written
index-for-index like the kernels so their geometry is validated on the CPU (tests/test_subpixel_math.py):

    nearest-2x-upsample -> ReflectionPad2d(2) -> Conv2d 5x5     (reference networks.py:256-257; 57 % of the decoder's MACs)

  ==  MAIN: one 3x3 convolution of the reflect-pad-1 SOURCE plane with 4*Cout output channels (the four output phases
            (py, px) folded into N; weights = sums of the 5x5 taps that read the same source pixel) + depth-to-space, valid for
            every source pixel that is not on the 1-pixel border ring of the source plane;
      RING: the 2-pixel border ring of the output (source-ring pixels x 4 phases), where the reference reflects in UPSAMPLED
            coordinates (row -1 reads S[0] but row -2 reads S[1]): recomputed by the plain 5x5 convolution on four thin strips
            of the exactly padded up-sampled plane (rows: [2n][6][2W+4]; columns, stored transposed: [2n][6][2H], convolved with
            the transposed filter).

9 taps per output pixel instead of 25 (+ ~10-17 % for the ring strips).  The backward pieces are the adjoints of the same maps:
space-to-depth of dY with the ring zeroed -> 3x3 data / weight gradients; ring rows / columns of dY -> 5x5 data / weight gradients
on the strips; strip input gradients gathered back onto the source pixels; phase weight gradients folded back onto the 5x5 taps."""
import torch
import torch.nn.functional as F

G = {0: ((0, 1), (2, 3), (4,)), 1: ((0,), (1, 2), (3, 4))}      # G[phase][source tap u] = the 5x5 taps reading source offset u-1


def reflect(i, L):
    if i < 0:
        i = -i
    if i >= L:
        i = 2 * (L - 1) - i
    return i


def src_of(y, L2):
    """source index read by coordinate y of the reflect-padded up-sampled axis of length L2 = 2*L"""
    return reflect(y, L2) // 2


def reference(S, W5, b=None):
    U = F.interpolate(S, scale_factor=2, mode="nearest")
    return F.conv2d(F.pad(U, (2, 2, 2, 2), mode="reflect"), W5, b)


def phase_weights(W5):
    """[4*Cout, Cin, 3, 3]: row (py*2+px)*Cout + co"""
    co, ci = W5.shape[:2]
    Wp = W5.new_zeros(4 * co, ci, 3, 3)
    for py in (0, 1):
        for px in (0, 1):
            ph = py * 2 + px
            for u in range(3):
                for v in range(3):
                    for a in G[py][u]:
                        for bb in G[px][v]:
                            Wp[ph * co:(ph + 1) * co, :, u, v] += W5[:, :, a, bb]
    return Wp


def fold_phase_grads(dWp, cout):
    """adjoint of phase_weights: dW5[co, ci, a, b] = sum over phases of dWp[phase row, ci, u(py, a), v(px, b)]"""
    ci = dWp.shape[1]
    dW5 = dWp.new_zeros(cout, ci, 5, 5)
    for py in (0, 1):
        for px in (0, 1):
            ph = py * 2 + px
            for u in range(3):
                for v in range(3):
                    for a in G[py][u]:
                        for bb in G[px][v]:
                            dW5[:, :, a, bb] += dWp[ph * cout:(ph + 1) * cout, :, u, v]
    return dW5


def row_strips(S):
    """[2n, C, 6, 2W+4]: rows -2..3 (top) and 2H-4..2H+1 (bottom) of the exactly padded up-sampled plane"""
    n, c, h, w = S.shape
    xs = [src_of(x - 2, 2 * w) for x in range(2 * w + 4)]
    top = [src_of(y - 2, 2 * h) for y in range(6)]
    bot = [src_of(2 * h - 4 + y, 2 * h) for y in range(6)]
    return torch.cat([S[:, :, top][:, :, :, xs], S[:, :, bot][:, :, :, xs]], 0), (top, bot, xs)


def col_strips(S):
    """[2n, C, 6, 2H] TRANSPOSED (strip row = up-sampled column -2..3 / 2W-4..2W+1, strip column = up-sampled row 0..2H-1)"""
    n, c, h, w = S.shape
    ys = [y // 2 for y in range(2 * h)]
    left = [src_of(x - 2, 2 * w) for x in range(6)]
    right = [src_of(2 * w - 4 + x, 2 * w) for x in range(6)]
    L = S[:, :, ys][:, :, :, left].transpose(2, 3)
    R = S[:, :, ys][:, :, :, right].transpose(2, 3)
    return torch.cat([L, R], 0), (left, right, ys)


def ring_mask(h, w, device=None):
    m = torch.zeros(h, w, dtype=torch.bool, device=device)
    m[0, :] = m[h - 1, :] = True
    m[:, 0] = m[:, w - 1] = True
    return m


def forward(S, W5, b=None):
    n, c, h, w = S.shape
    co = W5.shape[0]
    X = F.pad(S, (1, 1, 1, 1), mode="reflect")
    ym = F.conv2d(X, phase_weights(W5))                               # [n, 4co, h, w]
    y = S.new_zeros(n, co, 2 * h, 2 * w)
    keep = (~ring_mask(h, w)).to(S.dtype)
    for py in (0, 1):
        for px in (0, 1):
            ph = py * 2 + px
            y[:, :, py::2, px::2] = ym[:, ph * co:(ph + 1) * co] * keep       # depth-to-space, ring pixels not stored
    XR, _ = row_strips(S)
    yr = F.conv2d(XR, W5)                                             # [2n, co, 2, 2w]
    y[:, :, 0:2, :] = yr[:n]
    y[:, :, 2 * h - 2:, :] = yr[n:]
    XC, _ = col_strips(S)
    yc = F.conv2d(XC, W5.transpose(2, 3))                             # [2n, co, 2, 2h-4] (transposed: [x][y])
    y[:, :, 2:2 * h - 2, 0:2] = yc[:n].transpose(2, 3)
    y[:, :, 2:2 * h - 2, 2 * w - 2:] = yc[n:].transpose(2, 3)
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    return y


def backward(S, W5, Gy):
    """explicit adjoints: returns (gradient w.r.t. the reflect-pad-1 plane X [n, C, h+2, w+2], dW5)"""
    n, c, h, w = S.shape
    co = W5.shape[0]
    X = F.pad(S, (1, 1, 1, 1), mode="reflect")
    keep = (~ring_mask(h, w)).to(S.dtype)
    dYs = torch.cat([Gy[:, :, py::2, px::2] * keep for py in (0, 1) for px in (0, 1)], 1)      # space-to-depth, ring zeroed
    Wp = phase_weights(W5)
    gX = F.conv_transpose2d(dYs, Wp)                                  # [n, c, h+2, w+2]
    dW5 = fold_phase_grads(torch.nn.grad.conv2d_weight(X, Wp.shape, dYs), co)
    # ring rows
    XR, (top, bot, xs) = row_strips(S)
    GR = torch.cat([Gy[:, :, 0:2, :], Gy[:, :, 2 * h - 2:, :]], 0)
    dW5 = dW5 + torch.nn.grad.conv2d_weight(XR, W5.shape, GR)
    gXR = F.conv_transpose2d(GR, W5)                                  # [2n, c, 6, 2w+4]
    # ring columns (transposed strips, transposed filter)
    XC, (left, right, ys) = col_strips(S)
    GC = torch.cat([Gy[:, :, 2:2 * h - 2, 0:2].transpose(2, 3), Gy[:, :, 2:2 * h - 2, 2 * w - 2:].transpose(2, 3)], 0)
    W5T = W5.transpose(2, 3)
    dW5 = dW5 + torch.nn.grad.conv2d_weight(XC, W5T.shape, GC).transpose(2, 3)
    gXC = F.conv_transpose2d(GC, W5T)                                 # [2n, c, 6, 2h]
    # gather the strip input gradients back onto the source pixels (interior of the padded plane)
    for si, rows in enumerate((top, bot)):
        for yy, i in enumerate(rows):
            for xx, j in enumerate(xs):
                gX[:, :, i + 1, j + 1] += gXR[si * n:(si + 1) * n, :, yy, xx]
    for si, cols in enumerate((left, right)):
        for xx, j in enumerate(cols):
            for yy, i in enumerate(ys):
                gX[:, :, i + 1, j + 1] += gXC[si * n:(si + 1) * n, :, xx, yy]
    return gX, dW5


def fold_reflect1(gX):
    """adjoint of ReflectionPad2d(1): gradient w.r.t. the padded plane -> gradient w.r.t. the source"""
    g = gX[:, :, 1:-1, 1:-1].clone()
    g[:, :, 1, :] += gX[:, :, 0, 1:-1]
    g[:, :, -2, :] += gX[:, :, -1, 1:-1]
    g[:, :, :, 1] += gX[:, :, 1:-1, 0]
    g[:, :, :, -2] += gX[:, :, 1:-1, -1]
    g[:, :, 1, 1] += gX[:, :, 0, 0]
    g[:, :, 1, -2] += gX[:, :, 0, -1]
    g[:, :, -2, 1] += gX[:, :, -1, 0]
    g[:, :, -2, -2] += gX[:, :, -1, -1]
    return g


if __name__ == "__main__":
    torch.manual_seed(0)
    S = torch.randn(2, 6, 5, 7, dtype=torch.float64, requires_grad=True)
    W = torch.randn(4, 6, 5, 5, dtype=torch.float64, requires_grad=True)
    ref = reference(S, W)
    print("forward max err", float((forward(S, W) - ref).abs().max()))
    Gy = torch.randn_like(ref)
    gs, gw = torch.autograd.grad((ref * Gy).sum(), (S, W))
    gX, dW = backward(S.detach(), W.detach(), Gy)
    print("dS max err", float((fold_reflect1(gX) - gs).abs().max()), "dW max err", float((dW - gw).abs().max()))
