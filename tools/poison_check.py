"""Unwritten-halo check (VERDICT r1 "Robustness": activation planes are torch.empty, correctness depends on every producer
writing every element a consumer reads).  Runs one dis_update + gen_update + sample() of a small trainer twice - normally and with
engine.POISON on, where every plane that is allocated without a memset starts as bf16 NaNs - and compares every loss and every
gradient: a consumer that reads an element no producer wrote turns the losses non-finite.
usage (GPU box): python tools/poison_check.py [bf16|fp32x3] [dim] [--eager-too]"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch  # noqa: E402
import yaml  # noqa: E402
import engine as E  # noqa: E402
import trainer as T  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 32


def run(poison, cfgname, graphs):
    E.POISON = poison
    cfg = yaml.safe_load(open(os.path.join(ROOT, "acl-gan_b200", "configs", cfgname)))
    cfg["gen"].update(dim=dim, mlp_dim=64, n_res=2)
    cfg["dis"].update(dim=dim)
    cfg.update(display_size=2, precision=precision, cuda_graphs=graphs)
    torch.manual_seed(0)
    tr = T.aclgan_Trainer(copy.deepcopy(cfg)).cuda()
    torch.manual_seed(1)
    xa = (torch.rand(2, 3, 128, 128) * 2 - 1).cuda()
    xb = (torch.rand(2, 3, 128, 128) * 2 - 1).cuda()
    zs = [torch.randn(2, 8, 1, 1) for _ in range(6)]
    tr._noise = zs[:3]
    tr.dis_update(xa, xb, cfg)
    tr._noise = zs[3:]
    tr.gen_update(xa, xb, cfg)
    torch.cuda.synchronize()
    out = {k: float(getattr(tr, k)) for k in dir(tr) if k.startswith("loss_") and isinstance(getattr(tr, k), torch.Tensor)}
    for n in ("gen_AB", "gen_BA", "dis_A", "dis_B", "dis_2"):
        for k, p in getattr(tr, n).named_parameters():
            out["grad %s.%s" % (n, k)] = float(p.grad.double().norm())
    for i, t in enumerate(tr.sample(xa, xb)):
        out["sample %d" % i] = float(t.double().norm())
    E.POISON = False
    return out


bad = 0
for cfgname in ("male2female.yaml", "selfie2anime.yaml"):
    for graphs in ((1, 0) if "--eager-too" in sys.argv else (1,)):
        a, b = run(False, cfgname, graphs), run(True, cfgname, graphs)
        worst = (0.0, "")
        for k in a:
            if not (b[k] == b[k]) or abs(b[k]) == float("inf"):
                bad += 1
                print("NON-FINITE under poison:", cfgname, "graphs", graphs, k, a[k], b[k])
                continue
            e = abs(a[k] - b[k]) / max(abs(a[k]), 1e-12)
            worst = max(worst, (e, k))
        print("%s %s graphs=%d dim=%d: %d quantities, worst normal-vs-poisoned difference %.1e (%s)" % (
            cfgname, precision, graphs, dim, len(a), worst[0], worst[1]))
print("poison check:", "FAILED (%d non-finite quantities)" % bad if bad else "ok - no consumer reads an unwritten element")
sys.exit(1 if bad else 0)
