"""Populate oracle/_ref/ with the UNMODIFIED reference so it can travel to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/aclgan_oracle.py header).  The reference is a flat
script repository (no setup.py, nothing to compile), so "building" it is a verbatim file copy of its
hot-path modules from where they lie under /root/reference into the git-ignored directory oracle/_ref/
(listed in .gitignore, NOT in .gpurunignore: like the built .so it ships with the gpurun snapshot but
never enters the history).  Nothing is edited; oracle/ref_shim.py provides the CPU shim around it.

Used by: bench.py --impl reference (cpu_baseline.kind = "reference"), the optional eager-CUDA library bar
of bench.py, and tests/test_oracle.py (live cross-check).  Called from __graft_entry__.build() whenever
/root/reference is present.
"""
import filecmp
import os
import shutil
import sys

SRC = os.environ.get("ACLGAN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("networks.py", "trainer.py", "utils.py", "data.py", "train.py", "test.py", "configs/male2female.yaml",
         "LICENSE.md")


def make(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "trainer.py")):
        if verbose:
            print("oracle/make_ref.py: %s not present - keeping whatever oracle/_ref holds" % SRC)
        return os.path.isfile(os.path.join(DST, "trainer.py"))
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    if verbose:
        print("oracle/_ref: unmodified reference hot-path modules copied from", SRC)
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
