"""CPU oracle for the ACL-GAN convolutional training step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and there only as the checker.

This is a *functional restatement* (plain ``torch`` CPU ops, fp32 or fp64, driven by
``state_dict``-style ``{key: tensor}`` dictionaries) of the reference's hot path; it is
not the reference's module tree.  Each function cites the reference ``file:line`` whose
arithmetic it restates (paths relative to ``/root/reference``).

Parity pin: ``oracle/make_golden.py`` imports the UNMODIFIED reference (available only
in the build container) and writes ``tests/golden/*.pt``; ``tests/test_oracle.py``
checks this restatement against those fixtures (and, when ``/root/reference`` is
mounted, against the live reference).  The reference ships no golden vectors of its
own, so those generated fixtures are the pin.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------
# architecture description (what the reference builds from the YAML `gen` / `dis` dicts)
# --------------------------------------------------------------------------------------


def gen_layout(gen_cfg: dict, input_dim: int) -> dict:
    """Static description of AdaINGen (networks.py:112-133, 212-264)."""
    dim = gen_cfg["dim"]
    n_down = gen_cfg["n_downsample"]
    n_res = gen_cfg["n_res"]
    return dict(
        dim=dim, n_down=n_down, n_res=n_res, style_dim=gen_cfg["style_dim"],
        mlp_dim=gen_cfg["mlp_dim"], out_dim=gen_cfg["output_dim"], in_dim=input_dim,
        activ=gen_cfg["activ"], pad_type=gen_cfg["pad_type"],
        content_dim=dim * (2 ** n_down),
    )


def _act(x: torch.Tensor, kind: str) -> torch.Tensor:
    # networks.py:344-357 (only the branches a shipped config reaches)
    if kind == "relu":
        return torch.relu(x)
    if kind == "lrelu":
        return F.leaky_relu(x, 0.2)
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "none":
        return x
    raise AssertionError("Unsupported activation: {}".format(kind))


def _pad(x: torch.Tensor, p: int, pad_type: str) -> torch.Tensor:
    # networks.py:318-325
    if p == 0:
        return x
    if pad_type == "reflect":
        return F.pad(x, (p, p, p, p), mode="reflect")
    if pad_type == "replicate":
        return F.pad(x, (p, p, p, p), mode="replicate")
    if pad_type == "zero":
        return F.pad(x, (p, p, p, p))
    raise AssertionError("Unsupported padding type: {}".format(pad_type))


def instance_norm(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.InstanceNorm2d(affine=False): biased variance, eps inside the sqrt (networks.py:333)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def adain(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """AdaptiveInstanceNorm2d.forward (networks.py:490-503): batch_norm over (1, B*C, H, W)
    in training mode == per-(n, c) instance norm with per-(n, c) scale `weight`, shift `bias`."""
    b, c = x.shape[:2]
    return instance_norm(x, eps) * weight.view(b, c, 1, 1) + bias.view(b, c, 1, 1)


def layer_norm_munit(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Custom LayerNorm (networks.py:520-536): per-sample mean / UNBIASED std over C*H*W,
    eps added to the std, per-channel affine."""
    b = x.shape[0]
    flat = x.reshape(b, -1)
    mean = flat.mean(1).view(b, 1, 1, 1)
    std = flat.std(1).view(b, 1, 1, 1)
    y = (x - mean) / (std + eps)
    return y * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def conv_block(x, p: Params, prefix: str, stride: int, padding: int, norm: str, activ: str,
               pad_type: str = "reflect", adain_wb=None) -> torch.Tensor:
    """Conv2dBlock.forward (networks.py:365-371): pad -> conv(+bias) -> norm -> activation."""
    y = F.conv2d(_pad(x, padding, pad_type), p[prefix + "conv.weight"], p[prefix + "conv.bias"], stride=stride)
    if norm == "in":
        y = instance_norm(y)
    elif norm == "adain":
        y = adain(y, adain_wb[0], adain_wb[1])
    elif norm == "ln":
        y = layer_norm_munit(y, p[prefix + "norm.gamma"], p[prefix + "norm.beta"])
    elif norm != "none":
        raise AssertionError("Unsupported normalization: {}".format(norm))
    return _act(y, activ)


def res_block(x, p: Params, prefix: str, norm: str, activ: str, pad_type: str, adain_wb=None):
    """ResBlock.forward (networks.py:306-310); adain_wb = [(w0,b0),(w1,b1)] for the two convs."""
    a0 = adain_wb[0] if adain_wb is not None else None
    a1 = adain_wb[1] if adain_wb is not None else None
    y = conv_block(x, p, prefix + "model.0.", 1, 1, norm, activ, pad_type, a0)
    y = conv_block(y, p, prefix + "model.1.", 1, 1, norm, "none", pad_type, a1)
    return y + x


def content_encode(x, p: Params, L: dict, prefix: str = "enc_content.") -> torch.Tensor:
    """ContentEncoder (networks.py:230-245)."""
    a, pt = L["activ"], L["pad_type"]
    y = conv_block(x, p, prefix + "model.0.", 1, 3, "in", a, pt)
    for i in range(L["n_down"]):
        y = conv_block(y, p, prefix + "model.%d." % (1 + i), 2, 1, "in", a, pt)
    rb = prefix + "model.%d." % (1 + L["n_down"])
    for i in range(L["n_res"]):
        y = res_block(y, p, rb + "model.%d." % i, "in", a, pt)
    return y


def style_encode(x, p: Params, L: dict, prefix: str = "enc_style.") -> torch.Tensor:
    """StyleEncoder with n_downsample=4 (networks.py:126, 212-228)."""
    a, pt = L["activ"], L["pad_type"]
    y = conv_block(x, p, prefix + "model.0.", 1, 3, "none", a, pt)
    for i in range(4):
        y = conv_block(y, p, prefix + "model.%d." % (1 + i), 2, 1, "none", a, pt)
    y = y.mean(dim=(2, 3), keepdim=True)                       # AdaptiveAvgPool2d(1)
    return F.conv2d(y, p[prefix + "model.6.weight"], p[prefix + "model.6.bias"])


def mlp(style, p: Params, prefix: str = "mlp.") -> torch.Tensor:
    """MLP 8 -> mlp_dim -> mlp_dim -> n_adain (networks.py:280-292)."""
    h = style.reshape(style.shape[0], -1)
    h = torch.relu(F.linear(h, p[prefix + "model.0.fc.weight"], p[prefix + "model.0.fc.bias"]))
    h = torch.relu(F.linear(h, p[prefix + "model.1.fc.weight"], p[prefix + "model.1.fc.bias"]))
    return F.linear(h, p[prefix + "model.2.fc.weight"], p[prefix + "model.2.fc.bias"])


def split_adain_params(ap: torch.Tensor, n_res: int, c: int):
    """assign_adain_params (networks.py:154-163): per AdaIN module, in modules() order,
    first `c` columns -> bias (mean), next `c` -> weight (std)."""
    out = []
    for i in range(2 * n_res):
        blk = ap[:, i * 2 * c:(i + 1) * 2 * c]
        bias = blk[:, :c].contiguous().view(-1)
        weight = blk[:, c:2 * c].contiguous().view(-1)
        out.append((weight, bias))
    return out


def decode(content, style, p: Params, L: dict) -> torch.Tensor:
    """AdaINGen.decode + Decoder.forward (networks.py:147-152, 247-264)."""
    a, pt = L["activ"], L["pad_type"]
    c = L["content_dim"]
    wb = split_adain_params(mlp(style, p), L["n_res"], c)
    y = content
    for i in range(L["n_res"]):
        y = res_block(y, p, "dec.model.0.model.%d." % i, "adain", a, pt, wb[2 * i:2 * i + 2])
    idx = 1
    for i in range(L["n_down"]):
        y = F.interpolate(y, scale_factor=2, mode="nearest")   # nn.Upsample(scale_factor=2), networks.py:256
        y = conv_block(y, p, "dec.model.%d." % (idx + 1), 1, 2, "ln", a, pt)
        idx += 2
    return conv_block(y, p, "dec.model.%d." % idx, 1, 3, "none", "tanh", pt)


def gen_encode(x, p: Params, L: dict):
    """AdaINGen.encode (networks.py:141-145) -> (content, style)."""
    return content_encode(x, p, L), style_encode(x, p, L)


def avgpool_3s2(x):
    """nn.AvgPool2d(3, stride=2, padding=[1,1], count_include_pad=False) (networks.py:33)."""
    return F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)


def dis_forward(x, p: Params, D: dict) -> List[torch.Tensor]:
    """MsImageDis.forward (networks.py:38-57): per scale n_layer x (conv4x4 s2 + lrelu), conv1x1."""
    outs = []
    for s in range(D["num_scales"]):
        y = x
        for l in range(D["n_layer"]):
            y = conv_block(y, p, "cnns.%d.%d." % (s, l), 2, 1, D["norm"], D["activ"], D["pad_type"])
        n = D["n_layer"]
        y = F.conv2d(y, p["cnns.%d.%d.weight" % (s, n)], p["cnns.%d.%d.bias" % (s, n)])
        outs.append(y)
        x = avgpool_3s2(x)
    return outs


def lsgan(outs: Sequence[torch.Tensor], target: float, gan_type: str = "lsgan"):
    """sum over scales of the GAN term against a constant target: 'lsgan' mean((o - t)^2) (networks.py:67, 83, 98);
    'nsgan' F.binary_cross_entropy(F.sigmoid(o), t) (networks.py:68-72, 84-86, 99-103; the outer torch.mean there acts on the
    sum of two already-reduced scalars)."""
    loss = 0
    for o in outs:
        if gan_type == "lsgan":
            loss = loss + torch.mean((o - target) ** 2)
        elif gan_type == "nsgan":
            loss = loss + F.binary_cross_entropy(torch.sigmoid(o), torch.full_like(o, target))
        else:
            assert 0, "Unsupported GAN type: {}".format(gan_type)
    return loss


def calc_dis_loss(fake, real, p, D):   # networks.py:60-75
    gt = D.get("gan_type", "lsgan")
    return lsgan(dis_forward(fake, p, D), 0.0, gt) + lsgan(dis_forward(real, p, D), 1.0, gt)


def calc_gen_loss(fake, p, D):         # networks.py:77-89
    return lsgan(dis_forward(fake, p, D), 1.0, D.get("gan_type", "lsgan"))


def calc_gen_d2_loss(fake, real, p, D):  # networks.py:91-106
    gt = D.get("gan_type", "lsgan")
    return lsgan(dis_forward(fake, p, D), 1.0, gt) + lsgan(dis_forward(real, p, D), 0.0, gt)


def focus_translation(fg, bg, focus):
    """trainer.py:85-88: m = (focus+1)/2 broadcast to 3 channels; fg*m + bg*(1-m)."""
    m = ((focus + 1) / 2).repeat(1, 3, 1, 1)
    return fg * m + bg * (1 - m)


def focus_terms(focus_raw, cfg):
    """trainer.py:146-158 for one mask: returns (size_loss, digit_loss)."""
    m = (focus_raw + 1) / 2
    d, up, lo, eps = cfg["focus_delta"], cfg["focus_upper"], cfg["focus_lower"], cfg["focus_epsilon"]
    size = torch.relu(torch.sum(m - up)) ** 2 * d + torch.relu(torch.sum(lo - m)) ** 2 * d
    digit = torch.sum(1 / (torch.abs(m - 0.5) + eps))
    return size, digit


# --------------------------------------------------------------------------------------
# initialisation (trainer.py:19-52, utils.py:274-294) - must consume the CPU RNG in the
# reference's order so that seed-identical weights come out.
# --------------------------------------------------------------------------------------


def _gen_param_specs(L: dict):
    """(key, shape, kind) in nn.Module registration order of AdaINGen (== state_dict /
    optimizer order).  kind: 'w' conv/linear weight, 'b' bias, 'gamma', 'beta', 'buf_mean', 'buf_var'."""
    d, sd, md = L["dim"], L["style_dim"], L["mlp_dim"]
    specs = []

    def conv(prefix, cin, cout, k):
        specs.append((prefix + "conv.weight", (cout, cin, k, k), "w"))
        specs.append((prefix + "conv.bias", (cout,), "b"))

    # enc_style (networks.py:212-225)
    conv("enc_style.model.0.", L["in_dim"], d, 7)
    c = d
    for i in range(2):
        conv("enc_style.model.%d." % (1 + i), c, 2 * c, 4)
        c *= 2
    for i in range(2):
        conv("enc_style.model.%d." % (3 + i), c, c, 4)
    specs.append(("enc_style.model.6.weight", (sd, c, 1, 1), "w"))
    specs.append(("enc_style.model.6.bias", (sd,), "b"))
    # enc_content (networks.py:230-242)
    conv("enc_content.model.0.", L["in_dim"], d, 7)
    c = d
    for i in range(L["n_down"]):
        conv("enc_content.model.%d." % (1 + i), c, 2 * c, 4)
        c *= 2
    rb = "enc_content.model.%d." % (1 + L["n_down"])
    for i in range(L["n_res"]):
        conv(rb + "model.%d.model.0." % i, c, c, 3)
        conv(rb + "model.%d.model.1." % i, c, c, 3)
    # dec (networks.py:247-261) - norm registered before conv inside a Conv2dBlock (networks.py:328-363)
    for i in range(L["n_res"]):
        for j in range(2):
            pre = "dec.model.0.model.%d.model.%d." % (i, j)
            specs.append((pre + "norm.running_mean", (c,), "buf_mean"))
            specs.append((pre + "norm.running_var", (c,), "buf_var"))
            conv(pre, c, c, 3)
    idx = 1
    for i in range(L["n_down"]):
        pre = "dec.model.%d." % (idx + 1)
        specs.append((pre + "norm.gamma", (c // 2,), "gamma"))
        specs.append((pre + "norm.beta", (c // 2,), "beta"))
        conv(pre, c, c // 2, 5)
        c //= 2
        idx += 2
    conv("dec.model.%d." % idx, c, L["out_dim"], 7)
    # mlp (networks.py:280-289)
    n_adain = 2 * L["n_res"] * 2 * L["content_dim"]
    specs.append(("mlp.model.0.fc.weight", (md, sd), "w"))
    specs.append(("mlp.model.0.fc.bias", (md,), "b"))
    specs.append(("mlp.model.1.fc.weight", (md, md), "w"))
    specs.append(("mlp.model.1.fc.bias", (md,), "b"))
    specs.append(("mlp.model.2.fc.weight", (n_adain, md), "w"))
    specs.append(("mlp.model.2.fc.bias", (n_adain,), "b"))
    return specs


def _dis_param_specs(D: dict, input_dim: int):
    specs = []
    for s in range(D["num_scales"]):
        c_in, c = input_dim, D["dim"]
        for l in range(D["n_layer"]):
            specs.append(("cnns.%d.%d.conv.weight" % (s, l), (c, c_in, 4, 4), "w"))
            specs.append(("cnns.%d.%d.conv.bias" % (s, l), (c,), "b"))
            c_in, c = c, c * 2
        n = D["n_layer"]
        specs.append(("cnns.%d.%d.weight" % (s, n), (1, c_in, 1, 1), "w"))
        specs.append(("cnns.%d.%d.bias" % (s, n), (1,), "b"))
    return specs


def _default_conv_init(shape, is_linear=False):
    """nn.Conv2d / nn.Linear reset_parameters(): kaiming_uniform_(a=sqrt(5)) on the weight."""
    w = torch.empty(shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    return w


def _default_bias_init(wshape):
    fan_in = 1
    for s in wshape[1:]:
        fan_in *= s
    bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
    return torch.empty(wshape[0]).uniform_(-bound, bound)


def _construct(specs):
    """Replays the RNG draws of module construction: each Conv2d/Linear draws weight then
    bias (torch.nn.modules.conv/linear reset_parameters); LayerNorm gamma draws uniform_()
    (networks.py:517); buffers draw nothing."""
    p = {}
    i = 0
    while i < len(specs):
        key, shape, kind = specs[i]
        if kind == "w":
            p[key] = _default_conv_init(shape)
            bkey, bshape, _ = specs[i + 1]
            p[bkey] = _default_bias_init(shape)
            i += 2
            continue
        if kind == "gamma":
            p[key] = torch.empty(shape).uniform_()
        elif kind == "beta":
            p[key] = torch.zeros(shape)
        elif kind == "buf_mean":
            p[key] = torch.zeros(shape)
        elif kind == "buf_var":
            p[key] = torch.ones(shape)
        i += 1
    return p


def _construct_gen(L):
    """Module construction order differs from registration order inside a Conv2dBlock:
    the norm (LayerNorm gamma draw) is created BEFORE the conv (networks.py:328-363), which
    is also the registration order, so the spec order is the RNG order.  One exception:
    AdaINGen builds enc_style, enc_content, dec, mlp in that order (networks.py:126-133)."""
    return _construct(_gen_param_specs(L))


def _reinit(p: Params, specs, init_type: str):
    """weights_init (utils.py:274-294) applied in module-traversal (== spec) order."""
    for key, shape, kind in specs:
        if kind == "w":
            if init_type == "kaiming":
                torch.nn.init.kaiming_normal_(p[key], a=0, mode="fan_in")
            elif init_type == "gaussian":
                torch.nn.init.normal_(p[key], 0.0, 0.02)
            elif init_type == "xavier":
                torch.nn.init.xavier_normal_(p[key], gain=math.sqrt(2))
            elif init_type == "orthogonal":
                torch.nn.init.orthogonal_(p[key], gain=math.sqrt(2))
            elif init_type != "default":
                raise AssertionError("Unsupported initialization: {}".format(init_type))
        elif kind == "b":
            p[key].zero_()


class OracleTrainer:
    """State + the two update steps of aclgan_Trainer (trainer.py:14-59, 90-170, 247-293),
    restated over flat parameter dictionaries with a hand-written Adam (torch.optim.Adam
    semantics: L2 weight decay added to the gradient, bias-corrected moments)."""

    NETS = ("gen_AB", "gen_BA", "dis_A", "dis_B", "dis_2")

    def __init__(self, cfg: dict, dtype=torch.float32, construct: bool = True):
        self.cfg = cfg
        self.dtype = dtype
        self.L = gen_layout(cfg["gen"], cfg["input_dim_a"])
        self.D = dict(cfg["dis"])
        self.specs = {
            "gen_AB": _gen_param_specs(self.L), "gen_BA": _gen_param_specs(self.L),
            "dis_A": _dis_param_specs(self.D, cfg["input_dim_a"]),
            "dis_B": _dis_param_specs(self.D, cfg["input_dim_a"]),
            "dis_2": _dis_param_specs(self.D, cfg["input_dim_b"]),
        }
        self.style_dim = cfg["gen"]["style_dim"]
        self.alpha = cfg["alpha"]
        self.step = {"gen": 0, "dis": 0}
        self.lr = {"gen": cfg["lr"], "dis": cfg["lr"]}
        self.nets: Dict[str, Params] = {}
        if construct:
            # construction order trainer.py:19-23, then display noise :30-32, then init :49-52
            for n in self.NETS:
                self.nets[n] = _construct(self.specs[n])
            ds = int(cfg["display_size"])
            self.z_1 = torch.randn(ds, self.style_dim, 1, 1)
            self.z_2 = torch.randn(ds, self.style_dim, 1, 1)
            self.z_3 = torch.randn(ds, self.style_dim, 1, 1)
            for n in self.NETS:                       # self.apply(weights_init(init)): children in registration order
                _reinit(self.nets[n], self.specs[n], cfg["init"])
            for n in ("dis_A", "dis_B", "dis_2"):
                _reinit(self.nets[n], self.specs[n], "gaussian")
            self._finish()

    # -- state handling ---------------------------------------------------------------
    def _finish(self):
        for n in self.NETS:
            for k in list(self.nets[n].keys()):
                t = self.nets[n][k].to(self.dtype)
                trainable = not k.endswith("running_mean") and not k.endswith("running_var")
                self.nets[n][k] = t.requires_grad_(trainable)
        self.adam = {}
        for grp, names in (("gen", ("gen_AB", "gen_BA")), ("dis", ("dis_A", "dis_B", "dis_2"))):
            self.adam[grp] = {(n, k): (torch.zeros_like(v), torch.zeros_like(v))
                              for n in names for k, v in self.nets[n].items() if v.requires_grad}

    def load_state_dicts(self, sds: Dict[str, Params]):
        self.nets = {n: {k: v.detach().clone() for k, v in sds[n].items()} for n in self.NETS}
        self._finish()

    def state_dicts(self) -> Dict[str, Params]:
        return {n: {k: v.detach().clone() for k, v in self.nets[n].items()} for n in self.NETS}

    def _zero_grads(self, names):
        for n in names:
            for v in self.nets[n].values():
                v.grad = None

    def _adam(self, grp):
        """torch.optim.Adam.step with weight_decay (L2), betas=(beta1,beta2), eps 1e-8 (trainer.py:39-42)."""
        c = self.cfg
        b1, b2, wd, eps, lr = c["beta1"], c["beta2"], c["weight_decay"], 1e-8, self.lr[grp]
        self.step[grp] += 1
        t = self.step[grp]
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        with torch.no_grad():
            for (n, k), (m, v) in self.adam[grp].items():
                p = self.nets[n][k]
                if p.grad is None:
                    continue
                g = p.grad + wd * p
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
                p.addcdiv_(m, denom, value=-lr / bc1)

    # -- forward cycle shared by both updates (trainer.py:103-133 / 258-280) -------------
    def _cycle(self, x_a, x_b, z, need_recon: bool):
        G, L = self.nets, self.L
        focus = self.cfg["focus_loss"] > 0
        z_1, z_2, z_3 = z
        c_1 = content_encode(x_a, G["gen_AB"], L)
        c_2 = content_encode(x_a, G["gen_BA"], L)
        out = {}
        xB = decode(c_1, z_1, G["gen_AB"], L)
        xA = decode(c_2, self.alpha * z_2, G["gen_BA"], L)
        if focus:
            xB, out["focus_B"] = xB.split(3, 1)
            xA, out["focus_A"] = xA.split(3, 1)
            xB = focus_translation(xB, x_a, out["focus_B"])
            xA = focus_translation(xA, x_a, out["focus_A"])
        if need_recon:
            s_2 = style_encode(x_a, G["gen_BA"], L)
            c_4, s_4 = gen_encode(x_b, G["gen_AB"], L)
            rA = decode(c_2, s_2, G["gen_BA"], L)
            rB = decode(c_4, s_4, G["gen_AB"], L)
            if focus:
                rA = rA.split(3, 1)[0]
                rB = rB.split(3, 1)[0]
            out["x_A_recon"], out["x_B_recon"] = rA, rB
        c_3 = content_encode(xB, G["gen_BA"], L)
        xA2 = decode(c_3, z_3, G["gen_BA"], L)
        if focus:
            xA2, out["focus_A2"] = xA2.split(3, 1)
            xA2 = focus_translation(xA2, xB, out["focus_A2"])
        out.update(x_B_fake=xB, x_A_fake=xA, x_A2_fake=xA2,
                   pair1=torch.cat((x_a, xA), -3), pair2=torch.cat((x_a, xA2), -3))
        return out

    def draw_z(self, batch):
        """trainer.py:99-101 / 254-256: three CPU randn draws per update."""
        return [torch.randn(batch, self.style_dim, 1, 1).to(self.dtype) for _ in range(3)]

    def gen_losses(self, x_a, x_b, z):
        """trainer.py:90-165 without the optimizer step; returns (dict of loss_* scalars, tensors)."""
        c, N, D = self.cfg, self.nets, self.D
        t = self._cycle(x_a, x_b, z, need_recon=True)
        ls = {}
        ls["loss_gen_adv_A"] = (calc_gen_loss(t["x_A_fake"], N["dis_A"], D) +
                                calc_gen_loss(t["x_A2_fake"], N["dis_A"], D)) * 0.5
        ls["loss_gen_adv_B"] = calc_gen_loss(t["x_B_fake"], N["dis_B"], D)
        ls["loss_gen_adv_2"] = calc_gen_d2_loss(t["pair1"], t["pair2"], N["dis_2"], D)
        total = c["gan_w"] * ls["loss_gen_adv_A"] + c["gan_w"] * ls["loss_gen_adv_B"] + c["gan_cw"] * ls["loss_gen_adv_2"]
        if c["focus_loss"] > 0:
            acc = 0
            for tag in ("B", "A", "A2"):
                s, d = focus_terms(t["focus_" + tag], c)
                ls["loss_gen_focus_%s_size" % tag] = s
                ls["loss_gen_focus_%s_digit" % tag] = d
                acc = acc + s + d
            total = total + c["focus_loss"] * acc / x_a.size(2) / x_a.size(3) / x_a.size(0) / 3
        ls["loss_idt_A"] = torch.mean(torch.abs(t["x_A_recon"] - x_a))
        ls["loss_idt_B"] = torch.mean(torch.abs(t["x_B_recon"] - x_b))
        total = total + c["recon_x_w"] * ls["loss_idt_A"] + c["recon_x_w"] * ls["loss_idt_B"]
        ls["loss_gen_total"] = total
        return ls, t

    def dis_losses(self, x_a, x_b, z):
        """trainer.py:247-290.  The generators run without autograd: the reference lets
        autograd flow into them (no detach) but discards those grads at trainer.py:91."""
        c, N, D = self.cfg, self.nets, self.D
        with torch.no_grad():
            t = self._cycle(x_a, x_b, z, need_recon=False)
        ls = {}
        ls["loss_dis_A"] = (calc_dis_loss(t["x_A_fake"], x_a, N["dis_A"], D) +
                            calc_dis_loss(t["x_A2_fake"], x_a, N["dis_A"], D)) * 0.5
        ls["loss_dis_B"] = calc_dis_loss(t["x_B_fake"], x_b, N["dis_B"], D)
        ls["loss_dis_2"] = calc_dis_loss(t["pair1"], t["pair2"], N["dis_2"], D)
        ls["loss_dis_total"] = c["gan_w"] * ls["loss_dis_A"] + c["gan_w"] * ls["loss_dis_B"] + c["gan_cw"] * ls["loss_dis_2"]
        return ls, t

    def dis_update(self, x_a, x_b, z=None, step: bool = True):
        self._zero_grads(("dis_A", "dis_B", "dis_2"))
        z = z if z is not None else self.draw_z(x_a.size(0))
        ls, t = self.dis_losses(x_a, x_b, z)
        ls["loss_dis_total"].backward()
        if step:
            self._adam("dis")
        return {k: v.detach() for k, v in ls.items()}, t

    def gen_update(self, x_a, x_b, z=None, step: bool = True):
        self._zero_grads(("gen_AB", "gen_BA"))
        for n in ("dis_A", "dis_B", "dis_2"):           # D weight grads are discarded by the reference (trainer.py:248)
            for v in self.nets[n].values():
                v.requires_grad_(False)
        try:
            z = z if z is not None else self.draw_z(x_a.size(0))
            ls, t = self.gen_losses(x_a, x_b, z)
            ls["loss_gen_total"].backward()
        finally:
            for n in ("dis_A", "dis_B", "dis_2"):
                for v in self.nets[n].values():
                    v.requires_grad_(True)
        if step:
            self._adam("gen")
        return {k: v.detach() for k, v in ls.items()}, {k: v.detach() for k, v in t.items()}
