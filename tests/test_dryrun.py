"""Host-logic test without a GPU: runs the whole dis_update / gen_update orchestration (engine.py, networks.py,
trainer.py) on CPU tensors with the kernel LAUNCHES replaced by no-ops.  The plan builders (pure host C++) run
for real, so every convolution geometry of both updates is validated; numerical results are meaningless here
(the -m gpu tests check those)."""
import ctypes as C
import copy
import os

import pytest
import torch
import yaml

import aclgan_native as N
import engine as E
import networks
import trainer as T

LAUNCHERS = ["aclgan_igemm_launch", "aclgan_wgrad_launch", "aclgan_pack_img", "aclgan_norm_stats",
             "aclgan_norm_finalize", "aclgan_norm_apply", "aclgan_block_bwd_reduce", "aclgan_block_bwd_apply",
             "aclgan_norm_bwd_finalize", "aclgan_img_grad_pack", "aclgan_img_grad_unpack", "aclgan_pack_weight",
             "aclgan_adam_step", "aclgan_adam_advance", "aclgan_adam_units", "aclgan_norm_finalize_apply", "aclgan_norm_bwd_finalize_apply", "aclgan_zero", "aclgan_copy", "aclgan_axpby",
             "aclgan_avgpool3x3s2_fwd", "aclgan_avgpool3x3s2_bwd", "aclgan_style_head_fwd", "aclgan_style_head_bwd",
             "aclgan_mlp_fwd", "aclgan_mlp_bwd", "aclgan_dis_head_fwd", "aclgan_dis_head_bwd", "aclgan_focus_blend_fwd",
             "aclgan_focus_blend_bwd", "aclgan_loss_reduce", "aclgan_focus_grad", "aclgan_loss_combine", "aclgan_stats_to_bias", "aclgan_pack_nchw",
             "aclgan_unpack_plane", "aclgan_up_derive_weights", "aclgan_up_gather_strips", "aclgan_up_dy_pack",
             "aclgan_up_scatter_strips", "aclgan_up_fold_wgrad"]


class _Stub:
    def __init__(self, real, counts):
        self._real, self.counts = real, counts

    def __getattr__(self, name):
        if name in LAUNCHERS:
            def f(*a, **k):
                self.counts[name] = self.counts.get(name, 0) + 1
                return 0
            return f
        return getattr(self._real, name)


@pytest.fixture
def dry(monkeypatch):
    counts = {}
    real = N.lib()
    monkeypatch.setattr(N, "_lib", _Stub(real, counts))
    monkeypatch.setattr(E.Engine, "_check_device", lambda self: None)
    monkeypatch.setattr(networks, "_default_engine", {})
    orig = E.Engine.__init__
    monkeypatch.setattr(E.Engine, "__init__", lambda self, precision="bf16", device="cpu": orig(self, precision, "cpu"))
    monkeypatch.setattr(networks._EngineNet, "_ensure_bound", _ensure_bound_cpu)
    yield counts
    N._lib = real


def _ensure_bound_cpu(self):
    if self._eng is None:
        self.bind(networks.get_engine())
    if not self._bound:
        self._build_layers()
        self._bound = True
        if self._own_arena:
            self._arena.finalize()
            self.attach_grads()


def _cfg(name="male2female.yaml", tiny=True):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, "acl-gan_b200", "configs", name)))
    if tiny:
        cfg["gen"].update(dim=16, mlp_dim=32, n_res=2)
        cfg["dis"].update(dim=16)
        cfg["display_size"] = 2
    return cfg


@pytest.mark.parametrize("cfgname,precision", [("male2female.yaml", "bf16"), ("male2female.yaml", "fp32x3"),
                                               ("selfie2anime.yaml", "bf16")])
def test_updates_dryrun(dry, cfgname, precision):
    cfg = _cfg(cfgname)
    cfg["precision"] = precision
    cfg["cuda_graphs"] = 0
    torch.manual_seed(0)
    tr = T.aclgan_Trainer(copy.deepcopy(cfg))
    x_a = torch.rand(2, 3, 64, 64) * 2 - 1
    x_b = torch.rand(2, 3, 64, 64) * 2 - 1
    tr.dis_update(x_a, x_b, cfg)
    n_dis = dict(dry)
    tr.gen_update(x_a, x_b, cfg)
    # dis_update: 3 batched discriminator passes x 3 scales x 4 conv layers, one wgrad each
    assert n_dis["aclgan_igemm_launch"] > 50 and n_dis["aclgan_wgrad_launch"] == 3 * 3 * 4
    assert dry["aclgan_wgrad_launch"] > n_dis["aclgan_wgrad_launch"]
    for name in ("loss_dis_total", "loss_gen_total", "loss_idt_A", "loss_gen_adv_2"):
        assert getattr(tr, name).dim() == 0
    for p in list(tr.gen_AB.parameters()) + list(tr.dis_2.parameters()):
        assert p.grad is not None and p.grad.shape == p.shape
    out = tr.sample(x_a, x_b)
    assert len(out) == (9 if cfg["focus_loss"] > 0 else 7)


def test_full_width_plans_dryrun(dry):
    """all layer geometries of the real (dim 64) networks at 64x64"""
    cfg = _cfg(tiny=False)
    cfg["cuda_graphs"] = 0
    torch.manual_seed(0)
    tr = T.aclgan_Trainer(copy.deepcopy(cfg))
    x = torch.rand(1, 3, 64, 64) * 2 - 1
    tr.dis_update(x, x, cfg)
    tr.gen_update(x, x, cfg)


@pytest.mark.parametrize("h,w", [(72, 72), (96, 80), (68, 136)])
def test_ragged_image_sizes_dryrun(dry, h, w):
    """crop sizes that are multiples of 4 but not powers of two: every plan of both updates and of sample() builds
    (68 / 136: the up blocks' column strips are 64 / 128 pixels long, tests/test_plans.py::test_conv_wgrad_strip_plan_...)"""
    cfg = _cfg()
    cfg["cuda_graphs"] = 0
    torch.manual_seed(0)
    tr = T.aclgan_Trainer(copy.deepcopy(cfg))
    x = torch.rand(2, 3, h, w) * 2 - 1
    tr.dis_update(x, x, cfg)
    tr.gen_update(x, x, cfg)
    out = tr.sample(x, x)
    assert tuple(out[1].shape) == (2, 3, h, w) and tuple(out[2].shape) == (2, 1, h, w)


def test_too_small_image_is_a_shape_error(dry):
    """100 x 60: the third discriminator scale shrinks to 3 x 1, where the reference's ReflectionPad2d(1) raises as well"""
    cfg = _cfg()
    cfg["cuda_graphs"] = 0
    tr = T.aclgan_Trainer(copy.deepcopy(cfg))
    x = torch.rand(1, 3, 100, 60) * 2 - 1
    with pytest.raises(N.NativeError, match="ACLGAN_ERR_SHAPE"):
        tr.dis_update(x, x, cfg)


def test_nsgan_updates_dryrun(dry):
    """gan_type 'nsgan' (reference networks.py:68-72, 84-86, 99-103) takes the same path: the head kernel gets gan_kind = NSGAN"""
    cfg = _cfg("selfie2anime.yaml")
    cfg["dis"]["gan_type"] = "nsgan"
    cfg["cuda_graphs"] = 0
    tr = T.aclgan_Trainer(copy.deepcopy(cfg))
    x = torch.rand(2, 3, 64, 64) * 2 - 1
    tr.dis_update(x, x, cfg)
    tr.gen_update(x, x, cfg)
    assert dry["aclgan_dis_head_fwd"] == 18          # 3 discriminators x 3 scales per update
    cfg["dis"]["gan_type"] = "wgan"
    tr = T.aclgan_Trainer(copy.deepcopy(cfg))           # the reference, too, only fails where the loss is formed
    with pytest.raises(AssertionError, match="Unsupported GAN type"):
        tr.dis_update(x, x, cfg)
