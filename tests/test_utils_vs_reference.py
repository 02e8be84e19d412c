"""The host-side helpers of the drop-in `utils` module against the UNMODIFIED reference's own functions, run live on the CPU
(reference utils.py; needs /root/reference or oracle/_ref): what train.py / test.py call around the hot path must behave the
same - file names, file contents, selection rules, learning-rate schedule."""
import copy
import filecmp
import os
import re

import pytest
import torch

import ref_shim
import utils

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference sources not available")


@pytest.fixture(scope="module")
def ru():
    return ref_shim.import_reference()[2]


@pytest.mark.parametrize("iterations", [-1, 4])
def test_step_scheduler_sequence(ru, iterations):
    hp = dict(lr_policy="step", step_size=3, gamma=0.5, lr=1e-3)
    seqs = []
    for U in (utils, ru):
        opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
        if iterations != -1:                        # a resumed optimizer.pt carries initial_lr (trainer.py:318-322)
            for g in opt.param_groups:
                g["initial_lr"] = 1e-3
        sch = U.get_scheduler(opt, hp, iterations)
        seq = []
        for _ in range(8):
            opt.step()
            sch.step()
            seq.append(opt.param_groups[0]["lr"])
        seqs.append(seq)
    assert seqs[0] == seqs[1]
    assert utils.get_scheduler(opt, dict(lr_policy="constant")) is None and ru.get_scheduler(opt, dict(lr_policy="constant")) is None


def test_model_list_picks_the_same_checkpoint(ru, tmp_path):
    d = str(tmp_path)
    for n in ("gen_00000002.pt", "gen_00000010.pt", "dis_00000010.pt", "optimizer.pt", "gen_00000004.pt", "notes.txt"):
        open(os.path.join(d, n), "w").close()
    for key in ("gen", "dis"):
        assert utils.get_model_list(d, key) == ru.get_model_list(d, key)
    assert utils.get_model_list(d + "/missing", "gen") is None and ru.get_model_list(d + "/missing", "gen") is None


def test_image_grids_byte_identical_and_html_equivalent(ru, tmp_path):
    torch.manual_seed(0)
    outs = [torch.rand(3, c, 16, 16) * 2 - 1 for c in (3, 3, 1, 3, 1, 3, 1, 3, 1)]       # the 9-tuple sample() returns (focus on)
    d1, d2 = str(tmp_path / "mine"), str(tmp_path / "ref")
    os.makedirs(d1)
    os.makedirs(d2)
    utils.write_2images(outs, 2, d1, "train_00000010")
    ru.write_2images(outs, 2, d2, "train_00000010")
    assert sorted(os.listdir(d1)) == sorted(os.listdir(d2)) == ["gen_a2b_train_00000010.jpg"]
    assert filecmp.cmp(os.path.join(d1, "gen_a2b_train_00000010.jpg"), os.path.join(d2, "gen_a2b_train_00000010.jpg"), shallow=False)
    utils.write_html(os.path.join(d1, "index.html"), 30, 10, "images")
    ru.write_html(os.path.join(d2, "index.html"), 30, 10, "images")
    a, b = (re.sub(r"\s+", "", open(os.path.join(d, "index.html")).read()) for d in (d1, d2))
    assert a == b                                   # same headings, links, widths; only the whitespace differs
    assert [os.path.relpath(p, d1) for p in utils.prepare_sub_folder(d1)] == [os.path.relpath(p, d2) for p in ru.prepare_sub_folder(d2)]


def test_config_and_old_checkpoint_conversion(ru):
    cfg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "acl-gan_b200", "configs", "male2female.yaml")
    with ref_shim.cpu_shim():                       # (the reference calls yaml.load without a Loader)
        assert utils.get_config(cfg) == ru.get_config(cfg)
    keys = ["enc_content.model.0.norm.running_mean", "enc_content.model.3.model.2.model.1.norm.running_var",
            "enc_content.model.3.model.4.model.1.norm.running_var", "enc_content.model.0.conv.weight",
            "enc_style.model.0.norm.running_var", "dec.model.0.model.0.model.1.norm.running_mean", "enc.model.0.norm.running_mean",
            "gen.enc_content.model.1.norm.running_var"]
    sd = {"a": {k: torch.zeros(1) for k in keys}, "b": {k: torch.ones(1) for k in keys[:3]}, "c": {}}
    for name in ("MUNIT", "aclgan", "UNIT"):
        mine = utils.pytorch03_to_pytorch04(copy.deepcopy(sd), name)
        ref = ru.pytorch03_to_pytorch04(copy.deepcopy(sd), name)
        assert {k: sorted(v) for k, v in mine.items()} == {k: sorted(v) for k, v in ref.items()}, name


def test_write_loss_selects_the_same_attributes(ru):
    class Obj:
        loss_a, grad_norm, nwd_x, other, _loss_private = 1.0, 2.0, 3.0, 4.0, 5.0

        def loss_fn(self):
            return 0

    class Writer:
        def __init__(self):
            self.rows = []

        def add_scalar(self, tag, value, step):
            self.rows.append((tag, value, step))

    w1, w2 = Writer(), Writer()
    utils.write_loss(6, Obj(), w1)
    ru.write_loss(6, Obj(), w2)
    assert w1.rows == w2.rows and ("loss_a", 1.0, 7) in w1.rows
