"""Host-logic scenarios of the drop-in surface on the CPU with the kernel launches stubbed out (the fixture of
tests/test_dryrun.py: real plan builders, real engine / tape / arena / optimizer-table code, no numerics): the call orders a user
of the reference's trainer.py can produce, beyond the one dis_update -> gen_update order the step tests run."""
import copy
import tempfile

import torch

import trainer as T
import utils
from test_dryrun import _cfg, dry  # noqa: F401  (fixture)


def _mk(cfgname="male2female.yaml"):
    cfg = _cfg(cfgname)
    cfg["cuda_graphs"] = 0
    return T.aclgan_Trainer(copy.deepcopy(cfg)), cfg


def _x(b, s=64):
    return torch.rand(b, 3, s, s) * 2 - 1


def test_sample_before_any_update(dry):  # noqa: F811
    tr, cfg = _mk()
    assert len(tr.sample(_x(2), _x(2))) == 9                  # train.py:45-48,83-95 may dump images before the first step
    tr.dis_update(_x(2), _x(2), cfg)
    tr.gen_update(_x(2), _x(2), cfg)
    assert len(tr.sample(_x(2), _x(2))) == 9 and tr.training


def test_batch_and_image_size_change_between_updates(dry):  # noqa: F811
    tr, cfg = _mk()
    for b, s in ((2, 64), (3, 64), (2, 64), (2, 96), (1, 64)):   # (a last, smaller batch of an epoch; another crop size)
        tr.dis_update(_x(b, s), _x(b, s), cfg)
        tr.gen_update(_x(b, s), _x(b, s), cfg)
        assert tr.loss_gen_total.dim() == 0


def test_standalone_encode_decode_around_trainer_binding(dry):  # noqa: F811
    tr, cfg = _mk()
    c, s = tr.gen_AB.encode(_x(2))                               # test.py:96 calls encode on a generator the trainer never stepped
    assert tuple(c.shape) == (2, 64, 16, 16) and tuple(s.shape) == (2, 8, 1, 1)
    assert tuple(tr.gen_AB.decode(c, s).shape) == (2, 4, 64, 64)
    tr.dis_update(_x(2), _x(2), cfg)                             # rebinds the networks to the trainer's arenas
    tr.gen_update(_x(2), _x(2), cfg)
    c, s = tr.gen_BA.encode(_x(2, 96))
    assert tuple(tr.gen_BA.decode(c, torch.randn(2, 8, 1, 1)).shape) == (2, 4, 96, 96)


def test_save_resume_fresh_and_same_object(dry):  # noqa: F811
    tr, cfg = _mk()
    tr.dis_update(_x(2), _x(2), cfg)
    tr.gen_update(_x(2), _x(2), cfg)
    with tempfile.TemporaryDirectory() as dn:
        tr.save(dn, 0)
        tr2, cfg2 = _mk()
        assert tr2.resume(dn, cfg2) == 1                         # trainer.py:301-322: before the first step (train.py:65)
        tr2.dis_update(_x(2), _x(2), cfg2)
        tr2.gen_update(_x(2), _x(2), cfg2)
        tr2.update_learning_rate()
        assert tr.resume(dn, cfg) == 1                           # ... and on an object that already stepped
        tr.dis_update(_x(2), _x(2), cfg)
        tr.gen_update(_x(2), _x(2), cfg)
        assert len(tr.gen_AB._dense_grads) == len(tr2.gen_AB._dense_grads)


def test_load_state_dict_mid_run_marks_packed_weights_dirty(dry):  # noqa: F811
    tr, cfg = _mk()
    tr.dis_update(_x(2), _x(2), cfg)
    n0 = dry.get("aclgan_pack_weight", 0)
    tr.gen_AB.load_state_dict(copy.deepcopy(tr.gen_AB.state_dict()))
    tr.gen_update(_x(2), _x(2), cfg)
    assert dry["aclgan_pack_weight"] > n0                        # the bf16 packings are re-derived from the new masters


def test_eval_sample_and_discriminator_tensor_api(dry):  # noqa: F811
    tr, cfg = _mk("selfie2anime.yaml")
    tr.eval()
    assert len(tr.sample(_x(2), _x(2))) == 7 and tr.training     # the reference's sample() ends with self.train() (trainer.py:227)
    assert len(tr.dis_A.forward(_x(2))) == 3
    assert tr.dis_A.calc_dis_loss(_x(2), _x(2)).dim() == 0
    pair = torch.cat((_x(2), _x(2)), 1)
    assert tr.dis_2.calc_gen_d2_loss(pair, pair).dim() == 0 and tr.dis_B.calc_gen_loss(_x(2)).dim() == 0


def test_loss_weights_change_between_updates_and_write_loss(dry):  # noqa: F811
    tr, cfg = _mk()
    tr.dis_update(_x(2), _x(2), cfg)
    tr.gen_update(_x(2), _x(2), cfg)
    cfg2 = dict(cfg, gan_w=2, recon_x_w=0.5)
    tr.dis_update(_x(2), _x(2), cfg2)
    tr.gen_update(_x(2), _x(2), cfg2)
    assert len(tr._lplans) == 4                                  # loss-combination matrices are cached per weight set

    class Writer:
        def __init__(self):
            self.tags = []

        def add_scalar(self, tag, value, step):
            self.tags.append(tag)

    w = Writer()
    utils.write_loss(0, tr, w)                                   # utils.py:190-194 scans the trainer's attributes
    assert {"loss_dis_total", "loss_gen_total", "loss_gen_focus_A2_digit", "loss_idt_B"} <= set(w.tags)
