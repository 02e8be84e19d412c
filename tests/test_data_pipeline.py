"""Input pipeline (reference utils.py:43-100, data.py:104-131): the host pipeline keeps the reference's
transform order; the device pipeline hands uint8 HWC images to `DeviceLoader`, whose `aclgan_augment_u8` kernel must reproduce
RandomHorizontalFlip + ToTensor + Normalize(0.5, 0.5) bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import utils


def _make_folder(tmp_path, n=6, h=40, w=52):
    from PIL import Image
    rng = np.random.RandomState(0)
    d = tmp_path / "imgs"
    d.mkdir()
    for i in range(n):
        Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8)).save(str(d / ("im%02d.png" % i)))
    return str(d)


def test_host_pipeline_matches_reference_transforms(tmp_path):
    folder = _make_folder(tmp_path)
    loader = utils.get_data_loader_folder(folder, 3, False, new_size=32, height=32, width=32, num_workers=0, gpu_augment=False)
    batches = list(loader)
    assert len(batches) == 2 and batches[0].shape == (3, 3, 32, 32) and batches[0].dtype == torch.float32
    assert float(batches[0].min()) >= -1.0 and float(batches[0].max()) <= 1.0
    # reference order for training: flip, resize, crop, ToTensor, Normalize (utils.py:83-100)
    names = [type(t).__name__ for t in utils._transform_list(True, 32, 32, 32, True, False).transforms]
    assert names == ["RandomHorizontalFlip", "Resize", "RandomCrop", "ToTensor", "Normalize"]
    names = [type(t).__name__ for t in utils._transform_list(True, 32, 32, 32, True, True).transforms]
    assert names == ["Resize", "RandomCrop", "_ToUint8HWC"]


def test_uint8_view_shares_files_and_returns_hwc(tmp_path):
    from data import ImageFolder
    folder = _make_folder(tmp_path)
    ref = ImageFolder(folder, transform=utils._transform_list(False, None, 40, 52, False, False))
    raw = utils._Reformat(ref, utils._transform_list(False, None, 40, 52, False, True))
    assert len(raw) == len(ref) == 6
    u8, f32 = raw[2], ref[2]
    assert u8.dtype == torch.uint8 and tuple(u8.shape) == (40, 52, 3) and tuple(f32.shape) == (3, 40, 52)
    assert torch.equal(((u8.permute(2, 0, 1).float() / 255) - 0.5) / 0.5, f32)
    assert ref.transform is not raw.ds.transform          # the reference-format dataset keeps its own transform


@pytest.mark.gpu
def test_augment_kernel_bit_exact_and_device_loader(tmp_path):
    import aclgan_native as N
    torch.manual_seed(0)
    n, h, w = 5, 37, 64
    u8 = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8)
    flip = torch.tensor([0, 1, 1, 0, 1], dtype=torch.uint8)
    ref = ((u8.permute(0, 3, 1, 2).float().div(255)).sub(0.5)).div(0.5)
    ref = torch.stack([r.flip(-1) if f else r for r, f in zip(ref, flip)])
    out = torch.empty(n, 3, h, w, device="cuda")
    src, fl = u8.cuda(), flip.cuda()
    N.check(N.lib().aclgan_augment_u8(src.data_ptr(), fl.data_ptr(), out.data_ptr(), n, h, w,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "augment_u8")
    assert torch.equal(out.cpu(), ref)
    # end to end: the device loader (no flip / full-size crop) == the host pipeline on the same files
    folder = _make_folder(tmp_path)
    host = utils.get_data_loader_folder(folder, 2, False, new_size=None, height=40, width=52, num_workers=0, gpu_augment=False)
    dev = utils.get_data_loader_folder(folder, 2, False, new_size=None, height=40, width=52, num_workers=0, gpu_augment=True)
    assert isinstance(dev, utils.DeviceLoader) and len(dev) == len(host) == 3
    assert tuple(dev.dataset[0].shape) == (3, 40, 52)            # train.py:45-48 indexes .dataset for the display images
    got = [b for b in dev]
    want = [b for b in host]
    assert len(got) == 3
    for g, wv in zip(got, want):
        assert g.is_cuda and g.dtype == torch.float32
        assert torch.equal(g.cpu(), wv)
    # training mode: flips drawn in the main process; every sample is the image or its mirror
    devt = utils.get_data_loader_folder(folder, 6, True, new_size=None, height=40, width=52, num_workers=0, gpu_augment=True)
    (batch,) = list(devt)
    full = torch.stack([dev.dataset[i] for i in range(6)])
    for img in batch.cpu():
        assert any(torch.equal(img, f) or torch.equal(img, f.flip(-1)) for f in full)


def test_host_loaders_identical_to_reference_loaders(tmp_path):
    """get_all_data_loaders (reference utils.py:43-100) in folder and file-list mode against the UNMODIFIED reference's own
    loaders on the same seed: every train / test batch identical (same dataset order, shuffle, flip, resize, crop draws)."""
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference sources not available")
    from PIL import Image
    _, _, rutils = ref_shim.import_reference()
    rng = np.random.RandomState(0)
    tmp = str(tmp_path)
    for sub in ("trainA", "trainB", "testA", "testB"):
        os.makedirs(os.path.join(tmp, sub))
        for i in range(5):
            Image.fromarray(rng.randint(0, 256, (50 + 3 * i, 70, 3), dtype=np.uint8)).save(os.path.join(tmp, sub, "%d.png" % i))
        with open(os.path.join(tmp, sub + ".txt"), "w") as f:
            f.write("\n".join("%d.png" % i for i in (3, 1, 4, 0)) + "\n")
    conf = dict(batch_size=2, num_workers=0, new_size=40, crop_image_height=32, crop_image_width=36, data_kind="x",
                data_root=tmp, gpu_augment=0)
    conf_list = {k: v for k, v in conf.items() if k != "data_root"}
    for d, sub in (("train_a", "trainA"), ("train_b", "trainB"), ("test_a", "testA"), ("test_b", "testB")):
        conf_list["data_folder_" + d] = os.path.join(tmp, sub)
        conf_list["data_list_" + d] = os.path.join(tmp, sub + ".txt")
    for c in (conf, conf_list):
        torch.manual_seed(5)
        mine = [[b.clone() for b in loader] for loader in utils.get_all_data_loaders(c)]
        torch.manual_seed(5)
        ref = [[b.clone() for b in loader] for loader in rutils.get_all_data_loaders(c)]
        for i, (m, r) in enumerate(zip(mine, ref)):
            assert len(m) == len(r) == 2, (i, len(m), len(r))
            assert tuple(m[0].shape) == ((2, 3, 32, 36) if i < 2 else (2, 3, 40, 40))      # test loaders crop to new_size
            assert all(torch.equal(a, b) for a, b in zip(m, r)), ("loader %d differs from the reference's" % i)
