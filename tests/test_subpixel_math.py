"""CPU check of the math behind DESIGN.md section 8 item 4 (not yet a kernel): nearest-2x-upsample -> ReflectionPad2d(2)
-> conv5x5 (reference networks.py:256-257) == four 3x3 phase convolutions on the reflect-padded SOURCE plane plus a
correction that is non-zero only on the two outermost output rows / columns - forward and, through autograd, both gradients."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import subpixel_proto as SP  # noqa: E402


@pytest.mark.parametrize("n,cin,cout,h,w", [(2, 6, 4, 5, 7), (1, 3, 5, 3, 4), (1, 4, 4, 8, 3)])
def test_subpixel_decomposition_is_exact(n, cin, cout, h, w):
    torch.manual_seed(0)
    S = torch.randn(n, cin, h, w, dtype=torch.float64, requires_grad=True)
    W = torch.randn(cout, cin, 5, 5, dtype=torch.float64, requires_grad=True)
    b = torch.randn(cout, dtype=torch.float64)
    ref = SP.reference(S, W, b)
    got = SP.decomposed(S, W, b)
    assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12)
    G = torch.randn_like(ref)
    gs_ref, gw_ref = torch.autograd.grad((ref * G).sum(), (S, W))
    gs, gw = torch.autograd.grad((got * G).sum(), (S, W))
    assert torch.allclose(gs, gs_ref, rtol=1e-11, atol=1e-11) and torch.allclose(gw, gw_ref, rtol=1e-11, atol=1e-11)


def test_difference_plane_support_and_phase_mac_count():
    S = torch.randn(1, 2, 6, 6, dtype=torch.float64)
    D = SP.difference_plane(S)
    H2, W2 = D.shape[-2:]
    lines = torch.zeros(H2, W2, dtype=torch.bool)
    lines[1, :] = lines[H2 - 2, :] = True
    lines[:, 1] = lines[:, W2 - 2] = True
    assert float(D[:, :, ~lines].abs().max()) == 0.0 and float(D.abs().max()) > 0
    # 4 phases x 9 taps at source resolution = 9 taps per OUTPUT pixel instead of 25
    assert len(SP.phase_weights(torch.zeros(1, 1, 5, 5))) == 4
