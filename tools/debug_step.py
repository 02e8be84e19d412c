"""per-tensor gradient errors of gen_update (tiny golden case) vs the fp64 oracle, with / without CUDA graphs"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("acl-gan_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch  # noqa: E402
import aclgan_oracle as O  # noqa: E402
import trainer as T  # noqa: E402

g32 = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_fp32.pt"), weights_only=False)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else g32["batch"]
torch.manual_seed(1)
s = g32["size"]
x_a = (torch.rand(g32["batch"], 3, s, s) * 2 - 1)[:batch]
x_b = (torch.rand(g32["batch"], 3, s, s) * 2 - 1)[:batch]
torch.manual_seed(2)
zs = [torch.randn(g32["batch"], 8, 1, 1)[:batch] for _ in range(6)]
for graphs in (1,):
    cfg = copy.deepcopy(g32["cfg"])
    if os.environ.get("FOCUS", "1") == "0":
        cfg["focus_loss"] = 0
        cfg["gen"]["output_dim"] = 3
    ocfg = copy.deepcopy(cfg)
    cfg["precision"] = "fp32x3"
    cfg["cuda_graphs"] = graphs
    torch.manual_seed(0)
    tr = T.aclgan_Trainer(cfg)
    if "BIAS" in os.environ:       # move the focus mask away from the cusp of the digit loss at m = 0.5
        with torch.no_grad():
            for gnet in (tr.gen_AB, tr.gen_BA):
                list(gnet.dec.model)[-1].conv.bias[3] += float(os.environ["BIAS"])
    tr.cuda()
    sds = {n: {k: v.detach().cpu().clone() for k, v in getattr(tr, n).state_dict().items()} for n in O.OracleTrainer.NETS}
    ot = O.OracleTrainer(ocfg, dtype=torch.float64, construct=False)
    ot.load_state_dicts(sds)
    ols, _ = ot.gen_update(x_a.double(), x_b.double(), [z.double() for z in zs[3:]], step=False)
    ref = {(n, k): v.grad.clone() for n in ("gen_AB", "gen_BA") for k, v in ot.nets[n].items() if v.grad is not None}
    for opt in (tr.dis_opt, tr.gen_opt):
        for grp in opt.param_groups:
            grp["lr"] = 0.0
            grp["weight_decay"] = 0.0
    tr._noise = zs[3:]
    tr.gen_update(x_a.cuda(), x_b.cuda(), cfg)
    torch.cuda.synchronize()
    rows = []
    for n in ("gen_AB", "gen_BA"):
        for k, p in getattr(tr, n).named_parameters():
            r = ref[(n, k)]
            nr = float(r.norm())
            if nr < 1e-7:
                continue
            rows.append((float((p.grad.double().cpu() - r).norm()) / nr, n + "." + k))
    rows.sort(reverse=True)
    print("== batch %d graphs %d: median %.2e" % (batch, graphs, rows[len(rows) // 2][0]))
    for e, k in rows[:8]:
        print("   %.3e %s" % (e, k))
    for k in ("loss_gen_total", "loss_gen_adv_A", "loss_gen_adv_B", "loss_gen_adv_2", "loss_idt_A", "loss_idt_B"):
        print("   %s mine %.6f oracle %.6f" % (k, float(getattr(tr, k)), float(ols[k])))
