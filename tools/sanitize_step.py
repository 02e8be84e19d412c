"""One 64x64 dis_update + gen_update of the narrow ('tiny') networks for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_step.py [bf16|fp32x3]
Eager launches (no CUDA graphs) so that every kernel is instrumented individually; the side-stream schedule stays on."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acl-gan_b200"))
import torch  # noqa: E402
import yaml  # noqa: E402
import trainer as T  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
cfg = yaml.safe_load(open(os.path.join(ROOT, "acl-gan_b200", "configs", "male2female.yaml")))
DIM = int(os.environ.get("SAN_DIM", "16"))      # 32: res blocks of 128 channels (column-split epilogue groups, CTA pairs)
cfg["gen"].update(dim=DIM, mlp_dim=32, n_res=2)
cfg["dis"].update(dim=DIM)
cfg["display_size"] = 2
cfg["precision"] = prec
cfg["cuda_graphs"] = int(os.environ.get("SAN_GRAPHS", "0"))
torch.manual_seed(0)
tr = T.aclgan_Trainer(copy.deepcopy(cfg)).cuda()
xa = torch.rand(2, 3, 64, 64, device="cuda") * 2 - 1
xb = torch.rand(2, 3, 64, 64, device="cuda") * 2 - 1
for it in range(int(os.environ.get("SAN_ITERS", "1"))):
    tr.dis_update(xa, xb, cfg)
    tr.gen_update(xa, xb, cfg)
torch.cuda.synchronize()
out = tr.sample(xa, xb)
torch.cuda.synchronize()
print("sanitize_step ok (%s): loss_dis_total %.5f loss_gen_total %.5f" % (prec, float(tr.loss_dis_total), float(tr.loss_gen_total)))
