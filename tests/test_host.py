"""CPU-side checks: the C-ABI library loads and exports what include/aclgan_b200.h declares, the drop-in modules
reproduce the reference's class surface / state_dict / optimizer order / seeded initialisation, and the product
path refuses to run without a CUDA device (no CPU fallback)."""
import copy
import json
import os

import pytest
import torch
import yaml

import aclgan_native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg():
    return yaml.safe_load(open(os.path.join(ROOT, "acl-gan_b200", "configs", "male2female.yaml")))


def test_library_exports_every_declared_symbol():
    L = N.lib()
    syms = N.exported_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert L.aclgan_version() == 1
    assert b"sm_100a" in L.aclgan_build_info()


def test_ctypes_structs_match_header_layout():
    import ctypes as C
    # a plan built by the C side must read back consistently through the ctypes mirror
    d = N.ConvDesc(64, 64, 3, 1, 1, 0)
    a = N.Act()
    a.planes, a.n, a.h, a.w, a.c, a.pad = 1, 2, 8, 8, 64, 1
    a.data[0] = 0x10000
    o = N.OutSpec()
    o.N = 0
    p = N.IgemmPlan()
    w = (C.c_uint64 * 2)(0x20000, 0)
    assert N.lib().aclgan_plan_conv_fwd(C.byref(d), C.byref(a), w, C.byref(o), C.byref(p)) == 0
    assert (p.num_taps, p.cchunks, p.block_n, p.box_x * p.box_y * p.box_z) == (9, 1, 64, 128)
    assert p.out.N == 2 and p.out.H == 8 and p.b[0].base == 0x20000 and p.a[0][0].base == 0x10000
    assert [p.tap_bk[t] for t in range(9)] == [64 * t for t in range(9)]


def test_reference_surface():
    import trainer as T
    surf = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_surface.json")))
    torch.manual_seed(0)
    tr = T.aclgan_Trainer(_cfg())
    assert [n for n, _ in tr.named_children()] == surf["children"]
    for n in ("gen_AB", "dis_A", "dis_2"):
        mine = [[k, list(v.shape)] for k, v in getattr(tr, n).state_dict().items()]
        assert mine == surf[n], n
    name_of = {id(p): "%s.%s" % (n, k) for n in ("gen_AB", "gen_BA", "dis_A", "dis_B", "dis_2")
               for k, p in getattr(tr, n).named_parameters()}
    assert [name_of[id(p)] for p in tr.gen_opt.param_groups[0]["params"]] == surf["gen_opt_order"]
    assert [name_of[id(p)] for p in tr.dis_opt.param_groups[0]["params"]] == surf["dis_opt_order"]
    for attr in ("gen_update", "dis_update", "sample", "save", "resume", "update_learning_rate", "recon_criterion",
                 "focus_translation", "z_1", "z_2", "z_3", "style_dim", "alpha", "focus_lam", "dis_scheduler"):
        assert hasattr(tr, attr), attr


@pytest.mark.parametrize("case", ["tiny", "p0nf"])
def test_seeded_init_matches_reference(golden_dir, case):
    import trainer as T
    g = torch.load(os.path.join(golden_dir, "%s_fp32.pt" % case), weights_only=False)
    torch.manual_seed(0)
    tr = T.aclgan_Trainer(copy.deepcopy(g["cfg"]))
    for n, sig in g["init_sig"].items():
        mine = torch.stack([torch.stack([v.double().sum(), v.double().abs().sum(), (v.double() ** 2).sum()])
                            for v in getattr(tr, n).state_dict().values()]).sum(0)
        assert torch.allclose(mine, sig, rtol=1e-9, atol=1e-9), n


def test_checkpoint_roundtrip(tmp_path):
    import trainer as T
    cfg = _cfg()
    cfg["gen"].update(dim=8, mlp_dim=16, n_res=1)
    cfg["dis"].update(dim=8)
    tr = T.aclgan_Trainer(cfg)
    tr.save(str(tmp_path), 41)
    assert sorted(os.listdir(tmp_path)) == ["dis_00000042.pt", "gen_00000042.pt", "optimizer.pt"]
    tr2 = T.aclgan_Trainer(cfg)
    assert tr2.resume(str(tmp_path), cfg) == 42
    for (k, a), (_, b) in zip(tr.gen_BA.state_dict().items(), tr2.gen_BA.state_dict().items()):
        assert torch.equal(a, b), k
    assert set(torch.load(os.path.join(tmp_path, "gen_00000042.pt")).keys()) == {"AB", "BA"}
    assert set(torch.load(os.path.join(tmp_path, "dis_00000042.pt")).keys()) == {"A", "B", "2"}


def test_option_surface_errors():
    import networks
    with pytest.raises(AssertionError):
        networks.Conv2dBlock(3, 8, 3, 1, 1, norm="bogus", activation="relu", pad_type="reflect")
    with pytest.raises(AssertionError):
        networks.Conv2dBlock(3, 8, 3, 1, 1, norm="none", activation="relu", pad_type="bogus")
    with pytest.raises(NotImplementedError):
        networks.Conv2dBlock(3, 8, 3, 1, 1, norm="bn", activation="relu", pad_type="reflect")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import networks
    gen = networks.AdaINGen(3, _cfg()["gen"])
    with pytest.raises(N.NativeError):
        gen.encode(torch.zeros(1, 3, 64, 64))


def test_tape_runs_in_reverse_and_keeps_tags():
    """the hand-scheduled reverse pass: closures run last-in-first-out; tags only matter with side streams (GPU)"""
    import engine as E
    tape = E.Tape()
    order = []
    tape.push(lambda: order.append("a"))
    tape.tag = 3
    tape.push(lambda: order.append("b"))
    tape.push(lambda: order.append("c"))
    tape.tag = None
    tape.push(lambda: order.append("d"))
    assert [t for _, t in tape.ops] == [None, 3, 3, None]
    tape.backward()                 # no streams: plain sequential reverse order
    assert order == ["d", "c", "b", "a"] and not tape.ops
    off = E.Tape(enabled=False)
    off.push(lambda: order.append("x"))
    assert not off.ops


def test_sums_pool_counts_then_serves_slices():
    """statistics workspace: counting mode measures the update, the sized pool hands out disjoint zeroed slices"""
    import engine as E
    count = E.SumsPool("cpu")
    count.begin()
    a = count.take((2, 64, 2))
    b = count.take((1, 3, 2))
    assert a.shape == (2, 64, 2) and float(a.abs().sum()) == 0 and count.off == 256 + 6
    pool = E.SumsPool("cpu", count.off)
    pool.begin()
    a, b = pool.take((2, 64, 2)), pool.take((1, 3, 2))
    a.fill_(1.0)
    assert float(b.abs().sum()) == 0 and a.data_ptr() != b.data_ptr()
    c = pool.take((4, 64, 2))       # beyond the measured size: falls back to a fresh tensor, never overlaps
    assert c.shape == (4, 64, 2) and float(c.abs().sum()) == 0
    pool.begin()
    assert float(pool.take((2, 64, 2)).abs().sum()) == 0      # one memset per update


def test_dis_loss_attributes_are_plain_attributes_until_set():
    """loss_dis_* behave like the reference's attributes: absent before the first update (utils.write_loss reflects over
    the trainer), readable after; stored under names without 'loss' so the reflection does not log them twice"""
    import trainer as T
    cfg = _cfg()
    cfg["gen"].update(dim=8, mlp_dim=8, n_res=1)
    cfg["dis"].update(dim=8)
    tr = T.aclgan_Trainer(copy.deepcopy(cfg))
    assert not hasattr(tr, "loss_dis_total")
    tr.loss_dis_total = torch.tensor(1.5)
    assert float(tr.loss_dis_total) == 1.5
    assert not [k for k in vars(tr) if "loss" in k and k.startswith("_")]


def _adam_entry(shape, conv):
    """an aclgan_adam_tensor as trainer._adam_group fills it (no device pointers needed for the host-side unit count)"""
    import ctypes as C
    t = N.AdamTensor()
    d = [1] * (4 - len(shape)) + list(shape)
    for j in range(4):
        t.d[j] = d[j]
    if conv:
        desc = N.ConvDesc(d[1], d[0], d[2], 1, d[2] // 2, N.WINDOW_NONE)
        L = N.lib()
        idx = lambda a, b, c, e: L.aclgan_packed_weight_index(C.byref(desc), 0, a, b, c, e)
        base = idx(0, 0, 0, 0)
        strides = (idx(1, 0, 0, 0) - base, idx(0, 1, 0, 0) - base, (idx(0, 0, 1, 0) - base) if d[2] > 1 else 0,
                   (idx(0, 0, 0, 1) - base) if d[3] > 1 else 0)
        for j in range(4):
            t.gs[j] = strides[j]
        t.pk[0][0] = 1      # (any non-zero: "this tensor has packed planes")
    else:
        acc = 1
        for j in reversed(range(4)):
            t.gs[j] = acc
            acc *= d[j]
    return t


@pytest.mark.parametrize("shape,conv,units", [
    ((256, 256, 3, 3), True, 8 * 8),          # 32 x 32 x 9 tiles
    ((128, 64, 4, 4), True, 8 * 2),           # 16 taps: 16 output channels per tile (shared-memory budget)
    ((128, 256, 5, 5), True, 8 * 8),          # 25 taps: 16 x 32 tiles
    ((64, 3, 7, 7), True, 2 * 1),             # few input channels: one ci tile
    ((4, 64, 7, 7), True, 1 * 2),
    ((256,), False, 1),                       # bias: flat, 1024 elements per unit
    ((256, 8), False, 2),
    ((4096, 256), False, 1024),
])
def test_adam_work_units(shape, conv, units):
    """host side of the tiled Adam kernel (csrc/adam.cu): CTAs per table entry"""
    import ctypes as C
    t = _adam_entry(shape, conv)
    assert N.lib().aclgan_adam_units(C.byref(t)) == units
