"""Runs one of the reference's OWN driver scripts, byte for byte (oracle/_ref/train.py or test.py - the unmodified copies
oracle/make_ref.py makes), against THIS package: `python tests/ref_driver.py train.py --config ... --output_path ...`.

The driver's `from trainer import aclgan_Trainer` / `from utils import ...` (reference train.py:7-10, test.py:8-9) resolve to
acl-gan_b200/ because that directory is put first on sys.path (runpy.run_path does not add the script's own directory), and
`import tensorboardX` (train.py:19; not in this image) to tests/stubs/.  Test infrastructure only."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "acl-gan_b200")
STUBS = os.path.join(ROOT, "tests", "stubs")
REF = os.path.join(ROOT, "oracle", "_ref")


def driver_path(name):
    return os.path.join(REF, name)


def run(name, argv):
    """executes the driver in this process as __main__ with sys.argv = [name] + argv; returns its SystemExit code/message"""
    for p in (STUBS, PKG):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for m in ("trainer", "networks", "utils", "data"):          # must be this package's modules, not the reference's
        mod = sys.modules.get(m)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "")).startswith(PKG):
            del sys.modules[m]
    old = sys.argv
    sys.argv = [name] + list(argv)
    try:
        runpy.run_path(driver_path(name), run_name="__main__")
    except SystemExit as e:
        return e.code
    finally:
        sys.argv = old
    return None


if __name__ == "__main__":
    code = run(sys.argv[1], sys.argv[2:])
    if code not in (None, 0, "Finish training"):                # train.py:106 ends with sys.exit('Finish training')
        sys.exit(code)
