"""debug: per-block backward intermediates of the tiny decoder vs oracle autograd"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("acl-gan_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, torch.nn.functional as F
import aclgan_oracle as O, engine as E, trainer as T, networks as NW

g32 = torch.load(os.path.join(ROOT, "tests/golden/tiny_fp32.pt"), weights_only=False)
cfg = copy.deepcopy(g32["cfg"]); cfg["precision"] = "fp32x3"
torch.manual_seed(0)
tr = T.aclgan_Trainer(cfg).cuda(); tr._setup()
G = tr.gen_AB; L = O.gen_layout(cfg["gen"], 3)
p = {k: v.detach().cpu().double().requires_grad_(True) for k, v in G.state_dict().items() if not k.endswith(("running_mean", "running_var"))}
torch.manual_seed(7)
content = torch.randn(2, L["content_dim"], 16, 16); z = torch.randn(2, 8)
tr.eng.debug = {}
tape = E.Tape()
cplane = G._content_from_tensor(content.cuda()); cplane.requires_grad = True
style = E.ImgT(z.cuda(), requires_grad=True)
img = G.dec_fwd(tape, cplane, style)
gi = torch.randn(img.t.shape); img.add_grad(gi.cuda()); tape.backward(); torch.cuda.synchronize()

# oracle with retained intermediate grads (restating decode block by block)
hi = content.bfloat16(); c64 = (hi.double() + (content - hi.float()).bfloat16().double()).requires_grad_(True)
wb = O.split_adain_params(O.mlp(z.double(), p), L["n_res"], L["content_dim"])
acts = {}
def cb(x, prefix, stride, pad, norm, act, adain_wb=None):
    y = F.conv2d(O._pad(x, pad, "reflect"), p[prefix + "conv.weight"], p[prefix + "conv.bias"], stride=stride)
    y.retain_grad(); acts[prefix] = y
    if norm == "adain": y2 = O.adain(y, adain_wb[0], adain_wb[1])
    elif norm == "ln": y2 = O.layer_norm_munit(y, p[prefix + "norm.gamma"], p[prefix + "norm.beta"])
    else: y2 = y
    return O._act(y2, act)
y = c64
for i in range(L["n_res"]):
    h = cb(y, "dec.model.0.model.%d.model.0." % i, 1, 1, "adain", "relu", wb[2 * i])
    y = cb(h, "dec.model.0.model.%d.model.1." % i, 1, 1, "adain", "none", wb[2 * i + 1]) + y
idx = 1
for i in range(L["n_down"]):
    y = F.interpolate(y, scale_factor=2, mode="nearest")
    y = cb(y, "dec.model.%d." % (idx + 1), 1, 2, "ln", "relu"); idx += 2
out = cb(y, "dec.model.%d." % idx, 1, 3, "none", "tanh")
(out * gi.double()).sum().backward()
print("fwd err", float((img.t.double().cpu() - out).norm() / out.norm()))
name_of = {}
for name, blk in G.named_modules():
    if isinstance(blk, NW.Conv2dBlock) and blk._layer is not None:
        name_of[id(blk._layer)] = name + "."
for lid, recs in tr.eng.debug.items():
    nm = name_of[lid]
    ref = acts[nm].grad
    dy = recs[0]["dy"].double().cpu()
    lay = [b._layer for n_, b in G.named_modules() if isinstance(b, NW.Conv2dBlock) and b._layer is not None and id(b._layer) == lid][0]
    gw = lay.grad_views()[0].double().cpu()
    xp = recs[0]["xpad"].double().cpu().permute(0, 3, 1, 2)[:, :lay.cin]
    ref_w = torch.nn.grad.conv2d_weight(xp, tuple(lay.weight.shape), dy, stride=lay.stride)
    ora_w = p[nm + "conv.weight"].grad
    print("%-36s dY err %.1e | dW vs conv2d_weight(my x, my dY) %.1e | dW vs oracle %.1e | ref_w vs oracle %.1e" % (
        nm, float((dy - ref).norm() / ref.norm()), float((gw - ref_w).norm() / ref_w.norm()),
        float((gw - ora_w).norm() / ora_w.norm()), float((ref_w - ora_w).norm() / ora_w.norm())))
