"""Drop-in `utils` module (host-side helpers the reference's train.py / test.py import).

Same function names and call signatures as the reference's utils.py; the bodies are written
for this repo: `get_config` uses the safe YAML loader (the reference's bare `yaml.load` fails on PyYAML >= 6,
utils.py:105), image / HTML writers and the data-loader factory keep the reference's file naming so the output
folders look the same.  None of this is on the GPU hot path.
"""
import math
import os
import time

import torch
import torch.nn.init as init
import yaml
from torch.optim import lr_scheduler


# ------------------------------------------------------------------------------------------ config
def get_config(config):
    with open(config, "r") as stream:
        return yaml.safe_load(stream)


# ------------------------------------------------------------------------------------------ init / schedule
def weights_init(init_type="gaussian"):
    """utils.py:274-294: re-draws every Conv*/Linear* weight, zeroes biases; draws in module order"""
    draw = {
        "gaussian": lambda w: init.normal_(w, 0.0, 0.02),
        "xavier": lambda w: init.xavier_normal_(w, gain=math.sqrt(2)),
        "kaiming": lambda w: init.kaiming_normal_(w, a=0, mode="fan_in"),
        "orthogonal": lambda w: init.orthogonal_(w, gain=math.sqrt(2)),
        "default": lambda w: w,
    }
    assert init_type in draw, "Unsupported initialization: {}".format(init_type)

    def init_fun(m):
        name = m.__class__.__name__
        if (name.startswith("Conv") or name.startswith("Linear")) and hasattr(m, "weight"):
            draw[init_type](m.weight.data)
            if getattr(m, "bias", None) is not None:
                init.constant_(m.bias.data, 0.0)

    return init_fun


def get_scheduler(optimizer, hyperparameters, iterations=-1):
    policy = hyperparameters.get("lr_policy", "constant")
    if policy == "constant":
        return None
    if policy == "step":
        if iterations >= 0:
            for grp in optimizer.param_groups:       # StepLR(last_epoch != -1) expects initial_lr
                grp.setdefault("initial_lr", hyperparameters["lr"])
        return lr_scheduler.StepLR(optimizer, step_size=hyperparameters["step_size"],
                                   gamma=hyperparameters["gamma"], last_epoch=iterations)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % policy)


def get_model_list(dirname, key):
    """latest checkpoint whose file name contains `key` (lexicographic order == iteration order)"""
    if not os.path.exists(dirname):
        return None
    names = sorted(os.path.join(dirname, f) for f in os.listdir(dirname)
                   if os.path.isfile(os.path.join(dirname, f)) and key in f and ".pt" in f)
    return names[-1] if names else None


def pytorch03_to_pytorch04(state_dict_base, trainer_name):
    """reference utils.py:309-388, behaviour for behaviour: takes a checkpoint with the MUNIT-era keys 'a' / 'b'; for
    trainer_name 'MUNIT' the InstanceNorm running statistics PyTorch 0.3 stored in the content encoder (model.0-2 and the eight
    ResBlock convs of model.3) are dropped, for every other trainer name - 'aclgan' included - the dictionaries come back
    unchanged (the reference's second key list sits in a nested function that is never called)."""
    import re
    stale = re.compile(r"enc_content\.model\.([012]|3\.model\.[0-3]\.model\.[01])\.norm\.running_(mean|var)$")

    def core(sd):
        if trainer_name != "MUNIT":
            return dict(sd)
        return {k: v for k, v in sd.items() if not stale.search(k)}
    return {"a": core(state_dict_base["a"]), "b": core(state_dict_base["b"])}


def vgg_preprocess(batch):
    raise NotImplementedError("VGG perceptual loss is outside the B200 hot path (vgg_w = 0 in all configs)")


def load_vgg16(model_dir):
    raise NotImplementedError("VGG perceptual loss is outside the B200 hot path (vgg_w = 0 in all configs)")


# ------------------------------------------------------------------------------------------ logging
class Timer:
    def __init__(self, msg):
        self.msg, self.start_time = msg, None

    def __enter__(self):
        self.start_time = time.time()

    def __exit__(self, exc_type, exc_value, exc_tb):
        print(self.msg % (time.time() - self.start_time))


def write_loss(iterations, trainer, train_writer):
    for name in dir(trainer):
        if name.startswith("__") or callable(getattr(trainer, name)):
            continue
        if "loss" in name or "grad" in name or "nwd" in name:
            train_writer.add_scalar(name, getattr(trainer, name), iterations + 1)


def prepare_sub_folder(output_directory):
    image_directory = os.path.join(output_directory, "images")
    checkpoint_directory = os.path.join(output_directory, "checkpoints")
    for d in (image_directory, checkpoint_directory):
        if not os.path.exists(d):
            print("Creating directory: {}".format(d))
            os.makedirs(d)
    return checkpoint_directory, image_directory


def _grid(image_outputs, display_image_num, file_name):
    import torchvision.utils as vutils
    tensors = [im.expand(-1, 3, -1, -1) for im in image_outputs]
    data = torch.cat([t[:display_image_num] for t in tensors], 0)
    grid = vutils.make_grid(data.data, nrow=display_image_num, padding=0, normalize=True)
    vutils.save_image(grid, file_name, nrow=1)


def write_2images(image_outputs, display_image_num, image_directory, postfix):
    """the reference (utils.py:122-124) writes ALL rows sample() returns - they are one translation direction - into
    gen_a2b_<postfix>.jpg and no gen_b2a file (its write_html still links one)"""
    _grid(image_outputs, display_image_num, "%s/gen_a2b_%s.jpg" % (image_directory, postfix))


def write_html(filename, iterations, image_save_iterations, image_directory, all_size=1536):
    """the reference's index page (utils.py:139-171): a 30 s auto-refresh page with the two "current" grids, then for every
    saved iteration (newest first) the four test / train grids; same headings, links and widths"""
    rows = ["<h3>current</h3>"]

    def row(it, path):
        rows.append('<h3>iteration [%d] (%s)</h3>\n<p><a href="%s">\n<img src="%s" style="width:%dpx">\n</a><br>\n<p>' % (
            it, path.split("/")[-1], path, path, all_size))

    row(iterations, "%s/gen_a2b_train_current.jpg" % image_directory)
    row(iterations, "%s/gen_b2a_train_current.jpg" % image_directory)
    for j in range(iterations, image_save_iterations - 1, -1):
        if j % image_save_iterations == 0:
            for tag in ("a2b_test", "b2a_test", "a2b_train", "b2a_train"):
                row(j, "%s/gen_%s_%08d.jpg" % (image_directory, tag, j))
    with open(filename, "w") as f:
        f.write('<!DOCTYPE html>\n<html>\n<head>\n<title>Experiment name = %s</title>\n<meta http-equiv="refresh" content="30">\n'
                "</head>\n<body>\n%s\n</body></html>" % (os.path.basename(filename), "\n".join(rows)))


# ------------------------------------------------------------------------------------------ data
# Two pipelines behind the reference's factory signatures (utils.py:43-100):
#  * host (the reference's): workers decode, Resize, RandomCrop, RandomHorizontalFlip, ToTensor, Normalize -> fp32 NCHW batches;
#  * device (`gpu_augment`, default whenever CUDA is available): workers only decode / Resize / RandomCrop and hand over uint8
#    HWC images (4x fewer bytes to collate, pin and copy); `DeviceLoader` copies each pinned batch on a side stream one batch
#    ahead of the consumer and runs `aclgan_augment_u8` (flip + ToTensor + Normalize + NHWC -> NCHW, bit-identical to the
#    torchvision ops).  It yields CUDA tensors, so the `.cuda()` of train.py:67 is a no-op.
def _transform_list(train, new_size, height, width, crop, device_tail):
    from torchvision import transforms
    tf = []
    if new_size is not None:
        tf.append(transforms.Resize(new_size))
    if crop:
        tf.append(transforms.RandomCrop((height, width)))
    if device_tail:
        tf.append(_ToUint8HWC())
    else:
        if train:
            tf = [transforms.RandomHorizontalFlip()] + tf
        tf += [transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))]
    return transforms.Compose(tf)


class _ToUint8HWC:
    """PIL image -> uint8 tensor [H, W, 3] (no float conversion on the host)"""

    def __call__(self, img):
        import numpy as np
        return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())


class _Reformat(torch.utils.data.Dataset):
    """the same files as `base` through another transform (the uint8 view the device pipeline's workers read)"""

    def __init__(self, base, transform):
        import copy
        self.ds = copy.copy(base)
        self.ds.transform = transform

    def __getitem__(self, i):
        return self.ds[i]

    def __len__(self):
        return len(self.ds)


class DeviceLoader:
    """Iterates a DataLoader of uint8 NHWC batches and yields normalised fp32 NCHW CUDA tensors, one batch ahead: the pinned
    host batch is copied on a side stream while the previous step computes, then `aclgan_augment_u8` applies the per-sample
    horizontal flip (drawn from the main-process torch RNG, p = 0.5 as transforms.RandomHorizontalFlip), ToTensor and
    Normalize.  `.dataset` is the reference-format dataset (train.py:45-48 indexes it for the display images)."""

    def __init__(self, loader, dataset, train, device="cuda"):
        self.loader, self.dataset, self.train, self.device = loader, dataset, train, torch.device(device)
        self.batch_size = loader.batch_size
        self._stream = None

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        import ctypes as C
        import aclgan_native as N
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.device)
        n, h, w, c = batch.shape
        assert c == 3 and batch.dtype == torch.uint8
        flip = (torch.rand(n) < 0.5).to(torch.uint8) if self.train else None
        with torch.cuda.stream(self._stream):
            src = batch.to(self.device, non_blocking=True)
            fl = flip.to(self.device, non_blocking=True) if flip is not None else None
            out = torch.empty((n, 3, h, w), dtype=torch.float32, device=self.device)
            N.check(N.lib().aclgan_augment_u8(src.data_ptr(), fl.data_ptr() if fl is not None else 0, out.data_ptr(), n, h, w,
                                              C.c_void_p(self._stream.cuda_stream)), "augment_u8")
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return out, ev, (src, fl)

    def __iter__(self):
        it = iter(self.loader)
        nxt = None
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            out, ev, keep = nxt
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream().wait_event(ev)
            out.record_stream(torch.cuda.current_stream())
            yield out


def _make_loader(dataset_cls, ds_args, batch_size, train, new_size, height, width, num_workers, crop, gpu_augment):
    from torch.utils.data import DataLoader
    ref_ds = dataset_cls(*ds_args, transform=_transform_list(train, new_size, height, width, crop, False))
    if gpu_augment is None:
        gpu_augment = torch.cuda.is_available()
    if not gpu_augment:
        return DataLoader(dataset=ref_ds, batch_size=batch_size, shuffle=train, drop_last=True, num_workers=num_workers,
                          pin_memory=True)
    raw = _Reformat(ref_ds, _transform_list(train, new_size, height, width, crop, True))
    loader = DataLoader(dataset=raw, batch_size=batch_size, shuffle=train, drop_last=True, num_workers=num_workers,
                        pin_memory=True, persistent_workers=num_workers > 0, prefetch_factor=4 if num_workers > 0 else None)
    return DeviceLoader(loader, ref_ds, train)


def get_data_loader_folder(input_folder, batch_size, train, new_size=None, height=256, width=256, num_workers=4,
                           crop=True, gpu_augment=None):
    from data import ImageFolder
    return _make_loader(ImageFolder, (input_folder,), batch_size, train, new_size, height, width, num_workers, crop, gpu_augment)


def get_data_loader_list(root, file_list, batch_size, train, new_size=None, height=256, width=256, num_workers=4,
                         crop=True, gpu_augment=None):
    from data import ImageFilelist
    return _make_loader(ImageFilelist, (root, file_list), batch_size, train, new_size, height, width, num_workers, crop, gpu_augment)


def get_all_data_loaders(conf):
    bs, nw = conf["batch_size"], conf["num_workers"]
    size_a = conf.get("new_size", conf.get("new_size_a"))
    size_b = conf.get("new_size", conf.get("new_size_b"))
    h, w = conf["crop_image_height"], conf["crop_image_width"]
    ga = conf.get("gpu_augment")            # None: device pipeline whenever CUDA is available; 0 / 1 force
    ga = None if ga is None else bool(ga)
    # test loaders crop to new_size x new_size (reference utils.py:57-60), train loaders to the configured crop
    if "data_root" in conf:
        root = conf["data_root"]
        mk = lambda sub, train, size: get_data_loader_folder(os.path.join(root, sub), bs, train, size,
                                                             h if train else size, w if train else size, nw, True, ga)
        return mk("trainA", True, size_a), mk("trainB", True, size_b), mk("testA", False, size_a), mk("testB", False, size_b)
    mk = lambda folder, lst, train, size: get_data_loader_list(conf[folder], conf[lst], bs, train, size,
                                                               h if train else size, w if train else size, nw, True, ga)
    return (mk("data_folder_train_a", "data_list_train_a", True, size_a),
            mk("data_folder_train_b", "data_list_train_b", True, size_b),
            mk("data_folder_test_a", "data_list_test_a", False, size_a),
            mk("data_folder_test_b", "data_list_test_b", False, size_b))
