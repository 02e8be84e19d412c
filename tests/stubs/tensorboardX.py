"""Test-only stand-in for the `tensorboardX` package the reference's train.py imports (train.py:19, 59) and this image does
not have: a SummaryWriter that appends every scalar to <logdir>/scalars.tsv so a test can read what write_loss logged."""
import os


class SummaryWriter:
    def __init__(self, logdir=None, **_):
        self.logdir = logdir or "."
        os.makedirs(self.logdir, exist_ok=True)
        self._path = os.path.join(self.logdir, "scalars.tsv")

    def add_scalar(self, tag, value, global_step=None, **_):
        with open(self._path, "a") as f:
            f.write("%s\t%r\t%r\n" % (tag, float(value), global_step))

    def close(self):
        pass
