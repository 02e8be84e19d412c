"""Import the UNMODIFIED reference (/root/reference) on a CPU-only host.

TEST INFRASTRUCTURE ONLY (see oracle/aclgan_oracle.py header).  The reference hard-codes
``.cuda()`` (trainer.py:30-32, 65-67, 99-101, 254-256), calls ``yaml.load`` without a Loader
(utils.py:105) and imports ``tensorboardX`` (train.py:19).  This shim makes those work on a
CPU-only box WITHOUT touching the reference sources:

* ``torch.Tensor.cuda`` / ``nn.Module.cuda`` become identity while the shim is active,
* ``yaml.load`` defaults to ``SafeLoader``.

Usable where ``/root/reference`` exists (the build container) or where ``oracle/make_ref.py`` left its verbatim,
git-ignored copy in ``oracle/_ref/`` (it travels to the GPU box with the snapshot).  ``/root/reference`` itself is never
read on the GPU box; callers must check ``available()`` and degrade (skip / fall back to the port) when neither exists.
"""
import contextlib
import importlib
import os
import sys

import torch
import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_ref_dir():
    """/root/reference in the build container; on the GPU box the verbatim copy oracle/make_ref.py left in
    oracle/_ref/ (git-ignored, travels with the gpurun snapshot)"""
    for d in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(d, "trainer.py")):
            return d
    return "/root/reference"


REF_DIR = os.environ.get("ACLGAN_REFERENCE_DIR") or _default_ref_dir()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "trainer.py"))


@contextlib.contextmanager
def cpu_shim():
    saved = (torch.Tensor.cuda, torch.nn.Module.cuda, yaml.load)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _orig_load = saved[2]

    def _load(stream, Loader=None, **kw):
        return _orig_load(stream, Loader=Loader or yaml.SafeLoader, **kw)

    yaml.load = _load
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda, yaml.load = saved


def import_reference():
    """Returns (networks, trainer, utils) modules of the reference, imported under
    private names so they never shadow this repo's drop-in modules of the same name."""
    if not available():
        raise RuntimeError("reference sources not found at %s" % REF_DIR)
    mods = {}
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("networks", "trainer", "utils", "data")}
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_DIR)
    try:
        with cpu_shim():
            for name in ("utils", "networks", "trainer"):
                mods[name] = importlib.import_module(name)
    finally:
        sys.path[:] = saved_path
        for k in ("networks", "trainer", "utils", "data"):
            m = sys.modules.pop(k, None)
            if m is not None:
                sys.modules["aclgan_ref_" + k] = m
            if saved_mods[k] is not None:
                sys.modules[k] = saved_mods[k]
    return mods["networks"], mods["trainer"], mods["utils"]
