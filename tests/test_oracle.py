"""Pins oracle/aclgan_oracle.py (the CPU restatement) to the reference.

* against tests/golden/*.pt - produced by oracle/make_golden.py from the UNMODIFIED reference;
* against the live reference when /root/reference is mounted (build container only).
"""
import copy
import os

import pytest
import torch

import aclgan_oracle as O
import ref_shim

CASES = ["tiny", "p0", "p0nf", "nsgan"]


def _load(golden_dir, case, tag):
    return torch.load(os.path.join(golden_dir, "%s_%s.pt" % (case, tag)), weights_only=False)


def _inputs(g, dtype):
    torch.manual_seed(1)
    b, s = g["batch"], g["size"]
    x_a = (torch.rand(b, 3, s, s) * 2 - 1).to(dtype)
    x_b = (torch.rand(b, 3, s, s) * 2 - 1).to(dtype)
    torch.manual_seed(2)
    zs = [torch.randn(b, 8, 1, 1).to(dtype) for _ in range(6)]
    return x_a, x_b, zs


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("case", CASES)
def test_init_matches_reference_rng_stream(golden_dir, case):
    g = _load(golden_dir, case, "fp32")
    torch.manual_seed(0)
    tr = O.OracleTrainer(copy.deepcopy(g["cfg"]))
    for n, sig in g["init_sig"].items():
        mine = torch.stack([torch.stack([v.double().sum(), v.double().abs().sum(), (v.double() ** 2).sum()])
                            for v in tr.nets[n].values()]).sum(0)
        assert torch.allclose(mine, sig, rtol=1e-9, atol=1e-9), (n, mine, sig)


@pytest.mark.parametrize("case,tag", [(c, t) for c in CASES for t in ("fp32", "fp64")])
def test_updates_match_golden(golden_dir, case, tag):
    g = _load(golden_dir, case, tag)
    dtype = torch.float32 if tag == "fp32" else torch.float64
    tol = 2e-5 if tag == "fp32" else 1e-10
    torch.manual_seed(0)
    tr = O.OracleTrainer(copy.deepcopy(g["cfg"]), dtype=dtype)
    x_a, x_b, zs = _inputs(g, dtype)

    ls, t = tr.dis_update(x_a, x_b, zs[:3])
    for k, v in g["dis_losses"].items():
        assert abs(float(ls[k]) - float(v)) <= tol * abs(float(v)), (k, float(ls[k]), float(v))
    if tag == "fp32":
        for k in ("x_B_fake", "x_A_fake", "x_A2_fake"):
            assert _rel(t[k], g["dis_forward"][k]) < 1e-5, k
    gg = g["dis_grads"]
    gtol = 5e-3 if tag == "fp32" else 1e-8      # fp32 grads: the reference's own noise floor
    for i, key in enumerate(gg["keys"]):
        n, k = key.split(".", 1)
        gr = tr.nets[n][k].grad.double().reshape(-1)
        assert abs(float(gr.norm()) - float(gg["norm"][i])) <= gtol * float(gg["norm"][i]) + 1e-12, key
    ps = g["dis_params_after"]
    for i, key in enumerate(ps["keys"]):
        n, k = key.split(".", 1)
        p = tr.nets[n][k].detach().double()
        assert abs(float((p * p).sum()) - float(ps["sig"][i][2])) <= 1e-3 * float(ps["sig"][i][2]) + 1e-12, key

    ls, t = tr.gen_update(x_a, x_b, zs[3:])
    for k, v in g["gen_losses"].items():
        assert abs(float(ls[k]) - float(v)) <= 5 * tol * abs(float(v)) + 1e-12, (k, float(ls[k]), float(v))
    if tag == "fp32":
        for k in ("x_B_fake", "x_A_fake", "x_A2_fake", "x_A_recon", "x_B_recon"):
            assert _rel(t[k], g["gen_forward"][k]) < 1e-5, k
    gg = g["gen_grads"]
    worst = 0.0
    for i, key in enumerate(gg["keys"]):
        n, k = key.split(".", 1)
        gr = tr.nets[n][k].grad
        if gr is None:
            continue
        ref = float(gg["norm"][i])
        if ref < 1e-6:          # conv biases in front of IN/AdaIN: true gradient is exactly 0
            continue
        worst = max(worst, abs(float(gr.double().norm()) - ref) / ref)
    assert worst < (5e-2 if tag == "fp32" else 1e-7), worst


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not mounted (GPU box)")
def test_forward_matches_live_reference():
    nets_mod, trainer_mod, _ = ref_shim.import_reference()
    g_cfg = dict(dim=16, mlp_dim=32, style_dim=8, output_dim=4, activ="relu", n_downsample=2, n_res=2,
                 pad_type="reflect")
    d_cfg = dict(dim=16, norm="none", activ="lrelu", n_layer=3, gan_type="lsgan", num_scales=3, pad_type="reflect")
    torch.manual_seed(3)
    gen = nets_mod.AdaINGen(3, g_cfg).double()
    dis = nets_mod.MsImageDis(6, d_cfg).double()
    x = torch.rand(2, 3, 32, 32, dtype=torch.float64) * 2 - 1
    z = torch.randn(2, 8, 1, 1, dtype=torch.float64)
    L = O.gen_layout(g_cfg, 3)
    p = {k: v.detach() for k, v in gen.state_dict().items()}
    c_ref, s_ref = gen.encode(x)
    c, s = O.gen_encode(x, p, L)
    assert _rel(c, c_ref) < 1e-12 and _rel(s, s_ref) < 1e-12
    assert _rel(O.decode(c, z, p, L), gen.decode(c_ref, z)) < 1e-12
    x6 = torch.rand(2, 6, 32, 32, dtype=torch.float64)
    pd = {k: v.detach() for k, v in dis.state_dict().items()}
    for a, b in zip(O.dis_forward(x6, pd, d_cfg), dis.forward(x6)):
        assert _rel(a, b) < 1e-12
    assert abs(float(O.calc_dis_loss(x6, x6 * 0.5, pd, d_cfg)) - float(dis.calc_dis_loss(x6, x6 * 0.5))) < 1e-12
    # the dormant 'nsgan' option (networks.py:68-72, 84-86, 99-103): same weights, sigmoid + binary cross entropy
    dis.gan_type = "nsgan"
    n_cfg = dict(d_cfg, gan_type="nsgan")
    with ref_shim.cpu_shim():
        assert abs(float(O.calc_dis_loss(x6, x6 * 0.5, pd, n_cfg)) - float(dis.calc_dis_loss(x6, x6 * 0.5))) < 1e-12
        assert abs(float(O.calc_gen_loss(x6, pd, n_cfg)) - float(dis.calc_gen_loss(x6))) < 1e-12
        assert abs(float(O.calc_gen_d2_loss(x6, x6 * 0.5, pd, n_cfg)) - float(dis.calc_gen_d2_loss(x6, x6 * 0.5))) < 1e-12
