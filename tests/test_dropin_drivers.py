"""The drop-in claim of SURVEY.md 8(b), exercised end to end: the reference's UNMODIFIED train.py and test.py (oracle/_ref,
byte-for-byte copies made by oracle/make_ref.py) drive this package through tests/ref_driver.py -

    train.py  --config tiny.yaml                     3 iterations on an image-folder data set (D_update 1 / G_update 2 schedule,
                                                     write_loss every iteration, sample + write_2images + write_html, save)
    train.py  --config tiny.yaml --resume            resumes gen_00000002.pt / dis_ / optimizer.pt, trains on, saves again
    test.py   --checkpoint gen_00000004.pt ...       encode -> decode x num_style -> focus blend -> jpg files

* `test_drivers_dryrun` (CPU, `-m "not gpu"`): the same three invocations in-process with the kernel LAUNCHES stubbed out
  (tests/test_dryrun.py's fixture) and `.cuda()` mapped to the identity: every host-side line of the drivers, the loaders, the
  writers, save / resume and the inference API runs; numbers are meaningless.
* `test_drivers_on_gpu` (`-m gpu`): the real thing on cuda:0, CUDA graphs on (train.py in subprocesses; test.py in-process so
  the tensors it hands to save_image can be compared with the CPU oracle run from the same checkpoint, seed and input)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import yaml

import ref_driver
from test_dryrun import dry  # noqa: F401  (fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs_ref = pytest.mark.skipif(not os.path.isfile(ref_driver.driver_path("train.py")),
                               reason="oracle/_ref not populated (python oracle/make_ref.py where /root/reference exists)")


def _dataset(root, n=4, size=72):
    from PIL import Image
    rng = np.random.RandomState(0)
    for sub in ("trainA", "trainB", "testA", "testB"):
        os.makedirs(os.path.join(root, sub))
        for i in range(n):
            Image.fromarray(rng.randint(0, 256, (size, size, 3), dtype=np.uint8)).save(os.path.join(root, sub, "%d.png" % i))


def _config(tmp, max_iter, **extra):
    cfg = yaml.safe_load(open(os.path.join(ROOT, "acl-gan_b200", "configs", "male2female.yaml")))
    cfg["gen"].update(dim=16, mlp_dim=32, n_res=2)
    cfg["dis"].update(dim=16)
    cfg.update(data_root=os.path.join(tmp, "data"), new_size=64, crop_image_height=64, crop_image_width=64, num_workers=0,
               batch_size=2, max_iter=max_iter, log_iter=1, image_display_iter=1, image_save_iter=2, snapshot_save_iter=2,
               display_size=2)
    cfg.update(extra)
    path = os.path.join(tmp, "tiny.yaml")           # same file name both times: train.py derives the output folder from it
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return path, cfg


def _check_training_outputs(tmp, last_iter):
    out = os.path.join(tmp, "out", "outputs", "tiny")
    ck = os.path.join(out, "checkpoints")
    for name in ("gen_%08d.pt" % last_iter, "dis_%08d.pt" % last_iter, "optimizer.pt"):
        assert os.path.isfile(os.path.join(ck, name)), (name, os.listdir(ck))
    # (the reference's write_2images puts every row of sample() into the gen_a2b file and writes no gen_b2a file, utils.py:122-124)
    for name in ("gen_a2b_train_current.jpg", "gen_a2b_test_%08d.jpg" % last_iter, "gen_a2b_train_%08d.jpg" % last_iter):
        assert os.path.getsize(os.path.join(out, "images", name)) > 0
    assert os.path.isfile(os.path.join(out, "index.html")) and os.path.isfile(os.path.join(out, "config.yaml"))
    rows = [l.split("\t") for l in open(os.path.join(tmp, "out", "logs", "tiny", "scalars.tsv"))]
    tags = {r[0] for r in rows}
    # write_loss (reference utils.py:190-194) logs every trainer attribute with "loss" in its name
    for t in ("loss_dis_total", "loss_gen_total", "loss_gen_adv_2", "loss_idt_A", "loss_dis_2"):
        assert t in tags, sorted(tags)
    return ck, rows


def _test_py_outputs(folder, num_style, focus=True):
    names = ["input.jpg"] + ["output%03d.jpg" % j for j in range(num_style)]
    if focus:
        names += ["output%03d_mask.jpg" % j for j in range(num_style)] + ["output%03d_img.jpg" % j for j in range(num_style)]
    for n in names:
        assert os.path.getsize(os.path.join(folder, n)) > 0, n


@needs_ref
def test_drivers_dryrun(dry, tmp_path, monkeypatch):  # noqa: F811
    tmp = str(tmp_path)
    _dataset(os.path.join(tmp, "data"))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "manual_seed", lambda *a, **k: None)
    cfg_path, _ = _config(tmp, 3, cuda_graphs=0, gpu_augment=0)
    assert ref_driver.run("train.py", ["--config", cfg_path, "--output_path", os.path.join(tmp, "out")]) == "Finish training"
    ck, rows = _check_training_outputs(tmp, 2)
    assert {int(r[2]) for r in rows} == {1, 2, 3}
    n_dis = dry["aclgan_adam_step"]
    # 3 iterations of the D_update 1 / G_update 2 schedule (train.py:71-74): 3 dis_update + 2 gen_update = 5 Adam launches
    assert n_dis == 5, n_dis
    cfg_path, _ = _config(tmp, 4, cuda_graphs=0, gpu_augment=0)
    assert ref_driver.run("train.py", ["--config", cfg_path, "--output_path", os.path.join(tmp, "out"), "--resume"]) \
        == "Finish training"
    _check_training_outputs(tmp, 4)
    # iterations 2 and 3 after the resume: 2 dis_update + 1 gen_update (enumerate restarts at it = 0)
    assert dry["aclgan_adam_step"] == n_dis + 3, dry["aclgan_adam_step"]
    res = os.path.join(tmp, "res")
    assert ref_driver.run("test.py", ["--config", cfg_path, "--input", os.path.join(tmp, "data", "testA", "0.png"),
                                      "--output_folder", res, "--checkpoint", os.path.join(ck, "gen_00000004.pt"),
                                      "--num_style", "2"]) is None
    _test_py_outputs(res, 2)


def _run(args, tmp):
    env = dict(os.environ, PYTHONUNBUFFERED="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_driver.py")] + args, cwd=tmp, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    return p.stdout


@needs_ref
@pytest.mark.gpu
def test_drivers_on_gpu(tmp_path):
    import aclgan_oracle as O
    tmp = str(tmp_path)
    _dataset(os.path.join(tmp, "data"))
    cfg_path, cfg = _config(tmp, 3, precision="fp32x3")
    log = _run(["train.py", "--config", cfg_path, "--output_path", os.path.join(tmp, "out")], tmp)
    assert log.count("Elapsed time in update") == 3 and "Iteration: 00000003/00000003" in log, log[-2000:]
    ck, rows = _check_training_outputs(tmp, 2)
    vals = {(r[0], int(r[2])): float(r[1]) for r in rows}
    for it in (1, 2, 3):
        assert np.isfinite(vals[("loss_dis_total", it)]) and vals[("loss_dis_total", it)] > 0
    assert np.isfinite(vals[("loss_gen_total", 3)]) and vals[("loss_gen_total", 3)] > 0
    assert vals[("loss_dis_total", 1)] != vals[("loss_dis_total", 2)]          # the optimizer moved the discriminators
    step_before = torch.load(os.path.join(ck, "optimizer.pt"), map_location="cpu")["dis"]["state"][0]["step"]
    assert int(step_before) == 2, step_before                                # saved after iteration 2 (two dis_update)

    cfg_path, cfg = _config(tmp, 4, precision="fp32x3")
    log = _run(["train.py", "--config", cfg_path, "--output_path", os.path.join(tmp, "out"), "--resume"], tmp)
    assert "Resume from iteration 2" in log and log.count("Elapsed time in update") == 2, log[-2000:]
    _check_training_outputs(tmp, 4)
    opt = torch.load(os.path.join(ck, "optimizer.pt"), map_location="cpu")
    assert int(opt["dis"]["state"][0]["step"]) == 4 and int(opt["gen"]["state"][0]["step"]) == 2   # Adam counters continued
    g2 = torch.load(os.path.join(ck, "gen_00000002.pt"), map_location="cpu")["AB"]
    g4 = torch.load(os.path.join(ck, "gen_00000004.pt"), map_location="cpu")["AB"]
    moved = [k for k in g2 if g2[k].dtype.is_floating_point and not torch.equal(g2[k], g4[k])]
    assert len(moved) > len(g2) // 2, (len(moved), len(g2))

    # test.py in THIS process (its save_image calls recorded), then the same tensors from the CPU oracle: same checkpoint, input,
    # and style codes (test.py:38 seeds the CPU generator, the trainer's construction and :104's randn consume it in order)
    import torchvision.utils as vutils
    import trainer as T
    from PIL import Image
    from torchvision import transforms
    res = os.path.join(tmp, "res")
    inp = os.path.join(tmp, "data", "testA", "0.png")
    saved, orig_save = {}, vutils.save_image

    def recording_save(t, path, **kw):
        saved[os.path.basename(path)] = t.detach().float().cpu().clone()
        return orig_save(t, path, **kw)

    vutils.save_image = recording_save
    try:
        assert ref_driver.run("test.py", ["--config", cfg_path, "--input", inp, "--output_folder", res, "--checkpoint",
                                          os.path.join(ck, "gen_00000004.pt"), "--num_style", "2", "--seed", "7"]) is None
    finally:
        vutils.save_image = orig_save
    _test_py_outputs(res, 2)
    torch.manual_seed(7)
    T.aclgan_Trainer(yaml.safe_load(open(cfg_path)))
    z = torch.randn(2, 8, 1, 1)
    tf = transforms.Compose([transforms.Resize(64), transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
    x = tf(Image.open(inp).convert("RGB")).unsqueeze(0)
    L = O.gen_layout(cfg["gen"], cfg["input_dim_a"])
    p = {k: v.float() for k, v in g4.items()}
    with torch.no_grad():
        c, _ = O.gen_encode(x, p, L)
        for j in range(2):
            y = O.decode(c, z[j:j + 1], p, L)
            img, mask = y[:, :3], y[:, 3:4]
            want = {"output%03d.jpg" % j: (O.focus_translation(img, x, mask) + 1) / 2, "output%03d_img.jpg" % j: img,
                    "output%03d_mask.jpg" % j: mask.expand(-1, 3, -1, -1)}
            for name, w in want.items():
                err = float((saved[name] - w).norm() / w.norm())
                assert err < 1e-3, (name, err)
    assert float((saved["input.jpg"] - x).abs().max()) == 0.0
