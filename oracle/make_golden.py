"""Generate tests/golden/*.pt by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/{tiny,p0,p0nf,nsgan,p1_256,p2_256}_{fp32,fp64}.pt
                                            # (the two 256x256 cases take ~4 minutes each in fp64 on 8 cores)

Recipe: torch.manual_seed(0) -> aclgan_Trainer(cfg); manual_seed(1) ->
x_a, x_b = rand(B,3,H,H)*2-1; manual_seed(2) -> 6 style-noise draws (3 for dis_update,
3 for gen_update, fp32 randn exactly as trainer.py:254-256 / 99-101 draw them); then one
dis_update and one gen_update through the reference's own methods.  Recorded: init
checksums, every loss_* scalar, the generated images / masks, per-parameter gradient
norms + leading elements, and post-step parameter checksums.
"""
import copy
import os
import sys

import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def load_cfg(name):
    with open(os.path.join(ROOT, "acl-gan_b200", "configs", name)) as f:
        return yaml.safe_load(f)


def case_config(case):
    """p0 := configs/male2female.yaml at 64x64 bs=1 (BASELINE.json configs[0]);
    p0nf := same with the focus branch off (selfie2anime variant);
    tiny := narrow networks (dim 16) at 64x64 bs=2 for fast CPU/GPU parity sweeps;
    nsgan := the tiny networks, focus branch off, with dis.gan_type 'nsgan' (the dormant loss option of networks.py:68-72,
    84-86, 99-103)."""
    if case == "p0":
        cfg = load_cfg("male2female.yaml")
        return cfg, 1, 64
    if case == "p0nf":
        cfg = load_cfg("selfie2anime.yaml")
        return cfg, 1, 64
    if case == "tiny":
        cfg = load_cfg("male2female.yaml")
        cfg["gen"].update(dim=16, mlp_dim=32, n_res=2)
        cfg["dis"].update(dim=16)
        cfg["display_size"] = 2
        return cfg, 2, 64
    if case == "nsgan":
        cfg = load_cfg("selfie2anime.yaml")
        cfg["gen"].update(dim=16, mlp_dim=32, n_res=2)
        cfg["dis"].update(dim=16, gan_type="nsgan")
        cfg["display_size"] = 2
        return cfg, 2, 64
    if case == "p1_256":       # BASELINE.json configs[1] geometry (male2female 256x256), batch 2 to bound the CPU fp64 run
        return load_cfg("male2female.yaml"), 2, 256
    if case == "p2_256":       # BASELINE.json configs[2] geometry (selfie2anime := focus branch off, 256x256), batch 2
        return load_cfg("selfie2anime.yaml"), 2, 256
    raise ValueError(case)


def tensor_sig(t):
    t = t.detach().double().reshape(-1)
    return torch.stack([t.sum(), t.abs().sum(), (t * t).sum()])


def make_inputs(batch, size, dtype):
    torch.manual_seed(1)
    x_a = (torch.rand(batch, 3, size, size) * 2 - 1).to(dtype)
    x_b = (torch.rand(batch, 3, size, size) * 2 - 1).to(dtype)
    torch.manual_seed(2)
    zs = [torch.randn(batch, 8, 1, 1) for _ in range(6)]
    return x_a, x_b, zs


def run_case(case, dtype):
    nets_mod, trainer_mod, _ = ref_shim.import_reference()
    cfg, batch, size = case_config(case)
    with ref_shim.cpu_shim():
        torch.manual_seed(0)
        tr = trainer_mod.aclgan_Trainer(copy.deepcopy(cfg))
        init_sig = {n: torch.stack([tensor_sig(v) for v in getattr(tr, n).state_dict().values()]).sum(0)
                    for n in ("gen_AB", "gen_BA", "dis_A", "dis_B", "dis_2")}
        if dtype == torch.float64:
            tr.double()
        x_a, x_b, zs = make_inputs(batch, size, dtype)
        queue = [z.to(dtype) for z in zs]
        real_randn = torch.randn

        def fake_randn(*a, **k):          # hands the pre-drawn noise to trainer.py:254-256 / 99-101
            return queue.pop(0)

        out = {"case": case, "dtype": str(dtype), "batch": batch, "size": size, "cfg": cfg,
               "init_sig": init_sig}

        def forward_images(z):
            """Recompute the cycle's tensors with the reference's modules (no grad)."""
            with torch.no_grad():
                focus = cfg["focus_loss"] > 0
                c_1, _ = tr.gen_AB.encode(x_a)
                c_2, s_2 = tr.gen_BA.encode(x_a)
                c_4, s_4 = tr.gen_AB.encode(x_b)
                r = {}
                xB = tr.gen_AB.decode(c_1, z[0])
                xA = tr.gen_BA.decode(c_2, tr.alpha * z[1])
                rA = tr.gen_BA.decode(c_2, s_2)
                rB = tr.gen_AB.decode(c_4, s_4)
                if focus:
                    xB, r["focus_B"] = xB.split(3, 1)
                    xA, r["focus_A"] = xA.split(3, 1)
                    xB = tr.focus_translation(xB, x_a, r["focus_B"])
                    xA = tr.focus_translation(xA, x_a, r["focus_A"])
                    rA, rB = rA.split(3, 1)[0], rB.split(3, 1)[0]
                c_3, _ = tr.gen_BA.encode(xB)
                xA2 = tr.gen_BA.decode(c_3, z[2])
                if focus:
                    xA2, r["focus_A2"] = xA2.split(3, 1)
                    xA2 = tr.focus_translation(xA2, xB, r["focus_A2"])
                r.update(x_B_fake=xB, x_A_fake=xA, x_A2_fake=xA2, x_A_recon=rA, x_B_recon=rB,
                         content_1=c_1, style_2=s_2)
                r["dis_A_on_x_a"] = tr.dis_A.forward(x_a)
                r["dis_2_on_pair1"] = tr.dis_2.forward(torch.cat((x_a, xA), -3))
                return {k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in r.items()}

        def grads_of(names):
            """packed: keys [P], norm [P], head [P,8] (first 8 elements, zero padded), sig [P,3]"""
            keys, norms, heads, sigs = [], [], [], []
            for n in names:
                for k, p in getattr(tr, n).named_parameters():
                    if p.grad is None:
                        continue
                    gr = p.grad.detach().double().reshape(-1)
                    keys.append("%s.%s" % (n, k))
                    norms.append(gr.norm())
                    h = torch.zeros(8, dtype=torch.float64)
                    h[:min(8, gr.numel())] = gr[:8]
                    heads.append(h)
                    sigs.append(tensor_sig(gr))
            return dict(keys=keys, norm=torch.stack(norms), head=torch.stack(heads), sig=torch.stack(sigs))

        def snapshot(names):
            return {"%s.%s" % (n, k): p.detach().double().clone() for n in names for k, p in getattr(tr, n).named_parameters()}

        def params_sig(names, before):
            """post-step parameters: sig [P,3] of p_after, and of the UPDATE dp = p_after - p_before: dsig [P,3] (sum, abs-sum,
            sum of squares) + dhead [P,8] (first 8 elements) - an optimizer that does nothing, or steps the wrong way, shows
            up in the update even though it is invisible in sig (|dp| ~ lr = 1e-4 of |p|)"""
            keys = ["%s.%s" % (n, k) for n in names for k, p in getattr(tr, n).named_parameters()]
            ps = [p for n in names for k, p in getattr(tr, n).named_parameters()]
            sig = torch.stack([tensor_sig(p) for p in ps])
            dsig, dhead = [], []
            for key, p in zip(keys, ps):
                d = (p.detach().double() - before[key]).reshape(-1)
                dsig.append(tensor_sig(d))
                h = torch.zeros(8, dtype=torch.float64)
                h[:min(8, d.numel())] = d[:8]
                dhead.append(h)
            return dict(keys=keys, sig=sig, dsig=torch.stack(dsig), dhead=torch.stack(dhead))

        out["dis_forward"] = forward_images(queue[:3])
        before_d = snapshot(("dis_A", "dis_B", "dis_2"))
        torch.randn = fake_randn
        try:
            tr.dis_update(x_a, x_b, cfg)
        finally:
            torch.randn = real_randn
        out["dis_losses"] = {k: getattr(tr, k).detach().double() for k in
                             ("loss_dis_A", "loss_dis_B", "loss_dis_2", "loss_dis_total")}
        out["dis_grads"] = grads_of(("dis_A", "dis_B", "dis_2"))
        out["dis_params_after"] = params_sig(("dis_A", "dis_B", "dis_2"), before_d)

        out["gen_forward"] = forward_images(queue[:3])
        before_g = snapshot(("gen_AB", "gen_BA"))
        torch.randn = fake_randn
        try:
            tr.gen_update(x_a, x_b, cfg)
        finally:
            torch.randn = real_randn
        out["gen_losses"] = {k: v.detach().double() for k, v in vars(tr).items()
                             if k.startswith("loss_gen") or k.startswith("loss_idt")}
        out["gen_grads"] = grads_of(("gen_AB", "gen_BA"))
        out["gen_params_after"] = params_sig(("gen_AB", "gen_BA"), before_g)
        # keep the fixture small: images are 64x64; D maps / contents are stored in fp32
        for grp in ("dis_forward", "gen_forward"):
            if dtype == torch.float64:
                # fp64 fixtures pin losses / gradients / post-step parameters; the forward
                # tensors live in the fp32 fixture (fp32-vs-fp64 forward differs by ~1e-6)
                out[grp] = {k: v.float() for k, v in out[grp].items() if k in ("x_A2_fake",)}
            else:
                for k, v in out[grp].items():
                    out[grp][k] = [t.float() for t in v] if isinstance(v, list) else v.float()
            if size > 64:
                # 256x256 fixtures stay small: full-resolution tensors are replaced by their signature (sum, abs-sum, sum of
                # squares in fp64) and an 8x-strided sub-sample (the -m gpu test also compares full tensors with the live oracle)
                comp = {}
                for k, v in out[grp].items():
                    if isinstance(v, list) or v.dim() != 4 or v.numel() <= 32768:
                        comp[k] = v
                    else:
                        comp[k] = dict(sig=tensor_sig(v), sub=v[:, :, 3::8, 5::8].contiguous(), shape=tuple(v.shape))
                out[grp] = comp
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    cases = sys.argv[1:] or ["tiny", "p0", "p0nf", "nsgan", "p1_256", "p2_256"]
    for case in cases:
        for dtype, tag in ((torch.float32, "fp32"), (torch.float64, "fp64")):
            res = run_case(case, dtype)
            path = os.path.join(GOLDEN_DIR, "%s_%s.pt" % (case, tag))
            torch.save(res, path)
            print("wrote", path, os.path.getsize(path), "bytes;",
                  {k: float(v) for k, v in res["dis_losses"].items()},
                  {k: float(v) for k, v in res["gen_losses"].items() if "total" in k})


if __name__ == "__main__":
    main()
