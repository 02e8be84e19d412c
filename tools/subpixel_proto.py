"""Prototype (pure torch, fp64) of the sub-pixel decomposition of  nearest-2x-upsample -> ReflectionPad2d(2) -> conv5x5
(reference networks.py:256-257, the two up-blocks = 57 % of the decoder's MACs), DESIGN.md section 8 item 4.

    U = up2(S)                      (2H x 2W)
    Ut = reflect_pad(U, 2)          what the reference convolves            (rows -2 .. 2H+1)
    S1 = reflect_pad(S, 1)          (rows -1 .. H)
    V  = up2(S1)                    (rows -2 .. 2H+1): V[i] = S1[i // 2]
    D  = Ut - V                     non-zero ONLY on rows -1, 2H and columns -1, 2W:  D[-1] = S[0] - S[1], D[2H] = S[H-1] - S[H-2]

    conv5(Ut) = conv5(V) + conv5(D)
    conv5(V)[2i+a, 2j+b] = sum_{u,v in 0..2} Wab[u, v] * S1[i + u - 1, j + v - 1]        (four 3x3 phase kernels, 9/25 of the MACs)
        Wab[u, v] = sum_{kh in G_a(u)} sum_{kw in G_b(v)} W[kh, kw],   G_0 = ({0,1}, {2,3}, {4}),  G_1 = ({0}, {1,2}, {3,4})
    conv5(D) touches only the two outermost output rows / columns (thin correction GEMMs on the difference lines).

`decomposed()` evaluates the right-hand side; tests/test_subpixel_math.py checks it (and, through autograd, both gradients)
against the reference composition."""
import torch
import torch.nn.functional as F

G = {0: ((0, 1), (2, 3), (4,)), 1: ((0,), (1, 2), (3, 4))}


def reference(S, W, b=None):
    U = F.interpolate(S, scale_factor=2, mode="nearest")
    return F.conv2d(F.pad(U, (2, 2, 2, 2), mode="reflect"), W, b)


def phase_weights(W):
    """Wab [a][b] -> [Cout, Cin, 3, 3]"""
    out = {}
    for a in (0, 1):
        for b in (0, 1):
            rows = []
            for u in range(3):
                cols = []
                for v in range(3):
                    acc = 0
                    for kh in G[a][u]:
                        for kw in G[b][v]:
                            acc = acc + W[:, :, kh, kw]
                    cols.append(acc)
                rows.append(torch.stack(cols, -1))
            out[(a, b)] = torch.stack(rows, -2)
    return out


def difference_plane(S):
    """D = reflect_pad(up2(S), 2) - up2(reflect_pad(S, 1)): non-zero on 4 lines only"""
    Ut = F.pad(F.interpolate(S, scale_factor=2, mode="nearest"), (2, 2, 2, 2), mode="reflect")
    V = F.interpolate(F.pad(S, (1, 1, 1, 1), mode="reflect"), scale_factor=2, mode="nearest")
    return Ut - V


def decomposed(S, W, b=None):
    n, c, h, w = S.shape
    S1 = F.pad(S, (1, 1, 1, 1), mode="reflect")
    Wab = phase_weights(W)
    out = S.new_zeros(n, W.shape[0], 2 * h, 2 * w)
    for (a, bb), wk in Wab.items():
        out[:, :, a::2, bb::2] = F.conv2d(S1, wk)          # 3x3 on the source resolution
    D = difference_plane(S)
    corr = F.conv2d(D, W)                                   # (prototype: full conv; only the 2-pixel ring is non-zero)
    ring = torch.ones(2 * h, 2 * w, dtype=torch.bool)
    ring[2:-2, 2:-2] = False
    assert float(corr[:, :, ~ring].abs().max()) == 0.0
    out = out + corr
    if b is not None:
        out = out + b.view(1, -1, 1, 1)
    return out


if __name__ == "__main__":
    torch.manual_seed(0)
    S = torch.randn(2, 6, 5, 7, dtype=torch.float64)
    W = torch.randn(4, 6, 5, 5, dtype=torch.float64)
    print("max |decomposed - reference| =", float((decomposed(S, W) - reference(S, W)).abs().max()))
    D = difference_plane(S)
    H2, W2 = D.shape[-2:]
    lines = torch.zeros(H2, W2, dtype=torch.bool)
    lines[1, :] = lines[H2 - 2, :] = True           # padded rows -1 and 2H
    lines[:, 1] = lines[:, W2 - 2] = True           # padded columns -1 and 2W
    print("D is non-zero only on those four lines:", float(D[:, :, ~lines].abs().max()) == 0.0)
