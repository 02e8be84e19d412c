"""-m gpu: the inference path and checkpoint interchange (8f-2).

* AdaINGen.encode / decode (reference networks.py:141-152, driven by test.py:96-106) and aclgan_Trainer.sample
  (trainer.py:179-245) through the graph-captured batch-1 / batch-16 forward, against the CPU oracle with shared weights -
  including a SECOND call with different inputs (graph replay through the static buffers).
* a checkpoint triple WRITTEN BY THE UNMODIFIED REFERENCE (gen_*.pt / dis_*.pt / optimizer.pt, trainer.py:324-331) is resumed
  by the drop-in trainer on the GPU, trained one more iteration (Adam moments and step counter continue), saved, and resumed
  again by the reference."""
import copy
import os

import pytest
import torch

import aclgan_oracle as O
import trainer as T
from test_gpu_step import _build, _inputs, _load, _rel

pytestmark = pytest.mark.gpu


def _oracle_of(tr, cfg):
    ot = O.OracleTrainer(copy.deepcopy(cfg), construct=False)
    ot.load_state_dicts({n: {k: v.detach().cpu().clone() for k, v in getattr(tr, n).state_dict().items()} for n in ot.NETS})
    return ot


@pytest.mark.parametrize("case,batch,precision", [("p0", 1, "fp32x3"), ("p0nf", 16, "fp32x3"), ("p0", 16, "bf16")])
def test_encode_decode_vs_oracle(golden_dir, case, batch, precision):
    g = _load(golden_dir, case, "fp32")
    tr, cfg = _build(g, precision)
    ot = _oracle_of(tr, cfg)
    G, L = tr.gen_AB, ot.L
    p = ot.nets["gen_AB"]
    tol = 1e-3 if precision == "fp32x3" else 5e-2
    for rep in range(2):                    # rep 1 = graph replay with new inputs
        torch.manual_seed(10 + rep)
        x = torch.rand(batch, 3, 64, 64) * 2 - 1
        z = torch.randn(batch, 8, 1, 1)
        with torch.no_grad():
            c_ref, s_ref = O.gen_encode(x, p, L)
            y_ref = O.decode(c_ref, z, p, L)
            y2_ref = O.decode(c_ref, s_ref, p, L)
        c, s = G.encode(x.cuda())
        assert tuple(c.shape) == tuple(c_ref.shape) and tuple(s.shape) == tuple(s_ref.shape)
        assert _rel(c, c_ref) < tol and _rel(s, s_ref) < tol, (rep, _rel(c, c_ref), _rel(s, s_ref))
        y = G.decode(c, z.cuda())
        assert _rel(y, y_ref) < tol, (rep, "decode", _rel(y, y_ref))
        # decode from a content tensor that did NOT come from encode (test.py feeds arbitrary tensors): reference content
        y2 = G.decode(c_ref.cuda(), s_ref.cuda())
        assert _rel(y2, y2_ref) < tol, (rep, "decode(ref content)", _rel(y2, y2_ref))
        assert _rel(G.forward(x.cuda()), y2_ref) < tol
    assert len(G._igraphs) == 2, sorted(G._igraphs)         # one encode + one decode graph, replayed
    print("\n[encode/decode %s B=%d %s] content %.2e style %.2e decode %.2e" % (case, batch, precision, _rel(c, c_ref),
                                                                               _rel(s, s_ref), _rel(y, y_ref)))


@pytest.mark.parametrize("case", ["tiny", "p0nf"])
def test_sample_vs_oracle(golden_dir, case):
    """trainer.sample (trainer.py:179-245) incl. its focus blends / masks, fixed display noise z_1..z_3"""
    g = _load(golden_dir, case, "fp32")
    g = dict(g, cfg=dict(copy.deepcopy(g["cfg"]), display_size=2))
    g["init_sig"] = {} if case == "p0nf" else g["init_sig"]      # (display_size changes the RNG stream of the p0nf fixture)
    tr, cfg = _build(g, "fp32x3")
    ot = _oracle_of(tr, cfg)
    L, P = ot.L, ot.nets
    focus = cfg["focus_loss"] > 0
    for rep in range(2):
        torch.manual_seed(20 + rep)
        x_a = torch.rand(2, 3, 64, 64) * 2 - 1
        x_b = torch.rand(2, 3, 64, 64) * 2 - 1
        out = tr.sample(x_a.cuda(), x_b.cuda())
        z1, z2, z3 = (z.cpu() for z in (tr.z_1, tr.z_2, tr.z_3))
        ref = {k: [] for k in ("fa", "ma", "fb", "mb", "fa2", "ma2", "rec", "mrec")}
        with torch.no_grad():
            for i in range(2):
                a = x_a[i:i + 1]
                c_1, s_1 = O.gen_encode(a, P["gen_BA"], L)
                o = O.decode(c_1, z1[i:i + 1], P["gen_BA"], L)
                rec = O.decode(c_1, s_1, P["gen_BA"], L)
                ob = O.decode(O.content_encode(a, P["gen_AB"], L), z2[i:i + 1], P["gen_AB"], L)
                if focus:
                    fa = O.focus_translation(o[:, :3], a, o[:, 3:4])
                    fb = O.focus_translation(ob[:, :3], a, ob[:, 3:4])
                else:
                    fa, fb = o, ob
                o2 = O.decode(O.content_encode(fb, P["gen_BA"], L), z3[i:i + 1], P["gen_BA"], L)
                fa2 = O.focus_translation(o2[:, :3], fb, o2[:, 3:4]) if focus else o2
                for k, v in (("fa", fa), ("ma", o[:, 3:4]), ("fb", fb), ("mb", ob[:, 3:4]), ("fa2", fa2), ("ma2", o2[:, 3:4]),
                             ("rec", rec[:, :3] if focus else rec), ("mrec", rec[:, 3:4])):
                    ref[k].append(v)
            ref = {k: torch.cat(v) for k, v in ref.items()}
            if focus:
                want = (x_a, ref["fa"], ref["ma"], ref["fb"], ref["mb"], ref["fa2"], ref["ma2"], ref["rec"], ref["mrec"])
            else:
                c_4, s_4 = O.gen_encode(x_b, P["gen_AB"], L)
                rb = O.decode(c_4, s_4, P["gen_AB"], L)
                want = (x_a, ref["fa"], ref["fb"], ref["fa2"], ref["rec"], x_b, rb.repeat(2, 1, 1, 1))
        assert len(out) == len(want) == (9 if focus else 7)
        errs = []
        for i, (m, w) in enumerate(zip(out, want)):
            assert tuple(m.shape) == tuple(w.shape), (i, m.shape, w.shape)
            errs.append(_rel(m, w))
            assert errs[-1] < 1e-3, (rep, i, errs[-1])
    assert tr.training
    print("\n[sample %s] %d outputs, max rel err %.2e" % (case, len(out), max(errs)))


def test_reference_checkpoint_interchange(golden_dir, tmp_path):
    import ref_shim
    if not ref_shim.available():
        pytest.skip("needs the verbatim reference copy oracle/_ref (made by __graft_entry__.build() where /root/reference exists)")
    _, trainer_mod, _ = ref_shim.import_reference()
    g = _load(golden_dir, "tiny", "fp32")
    cfg = copy.deepcopy(g["cfg"])
    x_a, x_b, _ = _inputs(g)
    torch.manual_seed(5)
    zs = [torch.randn(g["batch"], 8, 1, 1) for _ in range(12)]
    real_randn = torch.randn
    d1, d2 = str(tmp_path / "ref"), str(tmp_path / "mine")
    os.makedirs(d1)
    os.makedirs(d2)

    def run_ref(rt, noise):
        queue = list(noise)
        torch.randn = lambda *a, **k: queue.pop(0)
        try:
            rt.dis_update(x_a, x_b, cfg)
            rt.gen_update(x_a, x_b, cfg)
        finally:
            torch.randn = real_randn

    with ref_shim.cpu_shim():
        torch.manual_seed(0)
        rt = trainer_mod.aclgan_Trainer(copy.deepcopy(cfg))
        with torch.no_grad():       # cusp-free operating point (tests/test_gpu_step.py docstring)
            for gnet in (rt.gen_AB, rt.gen_BA):
                list(gnet.dec.model)[-1].conv.bias[3] -= 1.5
        run_ref(rt, zs[:6])
        rt.save(d1, 0)                                   # checkpoint triple written by the reference
        run_ref(rt, zs[6:])
        ref_losses = {k: float(getattr(rt, k)) for k in ("loss_dis_total", "loss_gen_total", "loss_gen_adv_2", "loss_idt_A")}
        ref_params = {(n, k): p.detach().clone() for n in ("gen_AB", "gen_BA", "dis_A", "dis_B", "dis_2")
                      for k, p in getattr(rt, n).named_parameters()}
    assert sorted(os.listdir(d1)) == ["dis_00000001.pt", "gen_00000001.pt", "optimizer.pt"]

    torch.manual_seed(123)                               # different initial weights: everything must come from the checkpoint
    mine = T.aclgan_Trainer(dict(copy.deepcopy(cfg), precision="fp32x3")).cuda()
    assert mine.resume(d1, cfg) == 1
    mine._noise = zs[6:9]
    mine.dis_update(x_a.cuda(), x_b.cuda(), cfg)
    mine._noise = zs[9:12]
    mine.gen_update(x_a.cuda(), x_b.cuda(), cfg)
    torch.cuda.synchronize()
    for k, v in ref_losses.items():
        assert abs(float(getattr(mine, k)) - v) <= 1e-3 * abs(v), (k, float(getattr(mine, k)), v)
    assert float(mine._adam_gen["hyper"][6]) == 2.0 and float(mine._adam_dis["hyper"][6]) == 2.0
    # second-iteration update with the reference's Adam moments: a trainer that ignored exp_avg / exp_avg_sq / step would take
    # a first-step-sized (+-lr) update instead - every element off by O(lr)
    tot, big, sabs = 0, 0, 0.0
    lr = cfg["lr"]
    for (n, k), pr in ref_params.items():
        if k.endswith("conv.bias") and ("enc_content" in k or "dec.model.0." in k):
            continue                                     # norm-cancelled biases: round-off gradients
        d = ((dict(getattr(mine, n).named_parameters())[k].detach().cpu() - pr).abs() / lr).double()
        tot += d.numel()
        big += int((d > 0.25).sum())
        sabs += float(d.sum())
    assert big / tot < 2e-2 and sabs / tot < 5e-2, (big / tot, sabs / tot)
    mine.save(d2, 1)
    with ref_shim.cpu_shim():
        torch.manual_seed(7)
        rt2 = trainer_mod.aclgan_Trainer(copy.deepcopy(cfg))
        assert rt2.resume(d2, cfg) == 2                  # the reference reads the checkpoint the B200 trainer wrote
        for n in ("gen_AB", "gen_BA", "dis_A", "dis_B", "dis_2"):
            for (k, a), (_, b) in zip(getattr(rt2, n).state_dict().items(), getattr(mine, n).state_dict().items()):
                assert torch.equal(a.cpu(), b.cpu()), (n, k)
        st = rt2.gen_opt.state_dict()["state"]
        assert len(st) > 0 and all(float(v["step"]) == 2.0 for v in st.values())
    print("\n[reference checkpoint interchange] losses %s ; elements off by > 0.25 lr after the resumed iteration: %.2e, mean |dp| / lr "
          "%.2e" % ({k: "%.5f" % v for k, v in ref_losses.items()}, big / tot, sabs / tot))
