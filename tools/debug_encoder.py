"""debug: content-encoder backward intermediates vs oracle autograd (same flow as tests/test_gpu_step.py parts test)"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("acl-gan_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, torch.nn.functional as F
import aclgan_oracle as O, engine as E, trainer as T, networks as NW

DEBUG = int(os.environ.get("DBG", "1"))
g32 = torch.load(os.path.join(ROOT, "tests/golden/tiny_fp32.pt"), weights_only=False)
cfg = copy.deepcopy(g32["cfg"]); cfg["precision"] = "fp32x3"
torch.manual_seed(0)
tr = T.aclgan_Trainer(cfg).cuda(); tr._setup()
G = tr.gen_AB; L = O.gen_layout(cfg["gen"], 3)
p = {k: v.detach().cpu().double().requires_grad_(True) for k, v in G.state_dict().items() if not k.endswith(("running_mean", "running_var"))}
torch.manual_seed(1)
x_a = torch.rand(2, 3, 64, 64) * 2 - 1
if DEBUG:
    tr.eng.debug = {}
tape = E.Tape()
xin = E.ImgT(x_a.cuda(), requires_grad=True)
c = G.enc_content_fwd(tape, xin)
torch.manual_seed(5)
gc = torch.randn(c.n, c.c_valid, c.h, c.w)
gp = torch.zeros((c.n, c.h + 2, c.w + 2, c.c), dtype=tr.eng.prec.dtype, device="cuda")
gp[:, 1:-1, 1:-1, :c.c_valid] = gc.permute(0, 2, 3, 1).to(tr.eng.prec.dtype).cuda()
c.gp = gp
tape.backward(); torch.cuda.synchronize()

acts = {}
def cb(x, prefix, stride, pad, norm, act):
    y = F.conv2d(O._pad(x, pad, "reflect"), p[prefix + "conv.weight"], p[prefix + "conv.bias"], stride=stride)
    y.retain_grad(); acts[prefix] = y
    return O._act(O.instance_norm(y), act)
x64 = x_a.double().requires_grad_(True)
y = cb(x64, "enc_content.model.0.", 1, 3, "in", "relu")
y = cb(y, "enc_content.model.1.", 2, 1, "in", "relu")
y = cb(y, "enc_content.model.2.", 2, 1, "in", "relu")
for i in range(L["n_res"]):
    h = cb(y, "enc_content.model.3.model.%d.model.0." % i, 1, 1, "in", "relu")
    y = cb(h, "enc_content.model.3.model.%d.model.1." % i, 1, 1, "in", "none") + y
(y * gc.double()).sum().backward()
print("fwd err %.2e" % float((c.value_nchw().double().cpu() - y).norm() / y.norm()))
for name, blk in G.named_modules():
    if isinstance(blk, NW.Conv2dBlock) and blk._layer is not None and name.startswith("enc_content"):
        lay = blk._layer
        gw = lay.grad_views()[0].double().cpu()
        ora_w = p[name + ".conv.weight"].grad
        line = "%-40s dW vs oracle %.1e" % (name, float((gw - ora_w).norm() / ora_w.norm()))
        if DEBUG and id(lay) in tr.eng.debug:
            rec = tr.eng.debug[id(lay)][0]
            dy = rec["dy"].double().cpu(); ref = acts[name + "."].grad
            xp = rec["xpad"].double().cpu().permute(0, 3, 1, 2)
            line += " | dY err %.1e" % float((dy - ref).norm() / ref.norm())
            if lay.window == 0:
                ref_w = torch.nn.grad.conv2d_weight(xp[:, :lay.cin], tuple(lay.weight.shape), dy, stride=lay.stride)
                line += " | dW vs conv2d_weight(my x, my dY) %.1e" % float((gw - ref_w).norm() / ref_w.norm())
        print(line)
print("image grad err %.1e" % float((xin.grad.double().cpu() - x64.grad).norm() / x64.grad.norm()))
if DEBUG:
    # gp received by model.1's output plane == conv_transpose(dY2, W2) (gradient w.r.t. model.2's PADDED input)
    blocks = {n: b for n, b in G.named_modules() if isinstance(b, NW.Conv2dBlock)}
    for consumer, producer in (("enc_content.model.2", "enc_content.model.1"), ("enc_content.model.1", "enc_content.model.0")):
        lay2 = blocks[consumer]._layer
        dy2 = acts[consumer + "."].grad
        w2 = p[consumer + ".conv.weight"].detach()
        exp = F.conv_transpose2d(dy2, w2, stride=lay2.stride)
        got = tr.eng.debug[id(blocks[producer]._layer)][0]["gp"].double().cpu().permute(0, 3, 1, 2)[:, :exp.shape[1]]
        d = (got - exp)
        print("gp into %s: err %.2e ; per-row err (first image, channel-summed) max at rows %s" % (
            producer, float(d.norm() / exp.norm()),
            torch.topk(d[0].pow(2).sum((0, 2)), 4).indices.tolist()), "cols", torch.topk(d[0].pow(2).sum((0, 1)), 4).indices.tolist(),
            "per-image", [float(d[i].norm() / exp[i].norm()) for i in range(d.shape[0])])
        pad_ch = tr.eng.debug[id(blocks[producer]._layer)][0]["gp"][..., exp.shape[1]:]
        print("   padding channels abs max", float(pad_ch.abs().max()) if pad_ch.numel() else 0.0)

if DEBUG:
    name = "enc_content.model.1"
    rec = tr.eng.debug[id(blocks[name]._layer)][0]
    dy = rec["dy"].double().cpu(); ref = acts[name + "."].grad
    d = dy - ref
    print("dY1 per-image err", [float(d[i].norm() / ref[i].norm()) for i in range(2)])
    pc = d[0].pow(2).sum((1, 2)).sqrt() / ref[0].pow(2).sum((1, 2)).sqrt()
    print("dY1 image0 per-channel err", ["%.1e" % v for v in pc.tolist()])
    worst_c = int(pc.argmax())
    e = d[0, worst_c].abs()
    print("worst channel", worst_c, "abs err max %.3e mean %.3e ; ref abs mean %.3e" % (float(e.max()), float(e.mean()), float(ref[0, worst_c].abs().mean())))
    print("err rows profile", ["%.0e" % v for v in e.mean(1).tolist()])
    print("ratio dy/ref (median)", float((dy[0, worst_c] / ref[0, worst_c]).median()))
