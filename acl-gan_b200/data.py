"""Drop-in `data` module: image-folder / file-list datasets (host side, PIL).  Same class names as the
reference's data.py; not on the GPU hot path (the benchmark uses synthetic tensors)."""
import os

import torch.utils.data as data
from PIL import Image

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp")


def default_loader(path):
    return Image.open(path).convert("RGB")


def is_image_file(filename):
    return filename.lower().endswith(IMG_EXTENSIONS)


def make_dataset(directory):
    assert os.path.isdir(directory), "%s is not a valid directory" % directory
    found = []
    for root, _, names in sorted(os.walk(directory)):
        found += [os.path.join(root, n) for n in names if is_image_file(n)]
    return found


def default_flist_reader(flist):
    with open(flist, "r") as f:
        return [line.strip() for line in f if line.strip()]


class ImageFolder(data.Dataset):
    def __init__(self, root, transform=None, return_paths=False, loader=default_loader):
        imgs = sorted(make_dataset(root))
        if not imgs:
            raise RuntimeError("Found 0 images in: %s (supported: %s)" % (root, ",".join(IMG_EXTENSIONS)))
        self.root, self.imgs, self.transform, self.return_paths, self.loader = root, imgs, transform, return_paths, loader

    def __getitem__(self, index):
        path = self.imgs[index]
        img = self.loader(path)
        if self.transform is not None:
            img = self.transform(img)
        return (img, path) if self.return_paths else img

    def __len__(self):
        return len(self.imgs)


class ImageFilelist(data.Dataset):
    def __init__(self, root, flist, transform=None, flist_reader=default_flist_reader, loader=default_loader):
        self.root, self.imlist, self.transform, self.loader = root, flist_reader(flist), transform, loader

    def __getitem__(self, index):
        img = self.loader(os.path.join(self.root, self.imlist[index]))
        return self.transform(img) if self.transform is not None else img

    def __len__(self):
        return len(self.imlist)


class ImageLabelFilelist(data.Dataset):
    def __init__(self, root, flist, transform=None, flist_reader=default_flist_reader, loader=default_loader):
        self.root, self.transform, self.loader = root, transform, loader
        self.imlist = flist_reader(os.path.join(root, flist))
        self.classes = sorted({p.split("/")[0] for p in self.imlist})
        self.class_to_idx = {c: i for i, c in enumerate(self.classes)}
        self.imgs = [(p, self.class_to_idx[p.split("/")[0]]) for p in self.imlist]

    def __getitem__(self, index):
        path, label = self.imgs[index]
        img = self.loader(os.path.join(self.root, path))
        return (self.transform(img) if self.transform is not None else img), label

    def __len__(self):
        return len(self.imgs)
