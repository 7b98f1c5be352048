"""Writes tests/golden/image_pipeline.npz: outputs of the LIBRARIES the reference's image transforms call
(coarse_grained/fiber/transforms/transform.py:10-45) — Pillow `Image.resize(BICUBIC)` and torchvision
`Compose([Resize | RandomResizedCrop + RandomHorizontalFlip, ToTensor, Normalize])` — executed in the build container
on seeded synthetic images.  The oracle (oracle/image_oracle.py) and the CUDA path are checked against these bytes.

    python tools/make_golden_images.py        (Pillow 12.2.0, torchvision 0.26.0 at the time of writing)
"""
import os
import sys

import numpy as np
import PIL
import torch
import torchvision
from PIL import Image
from torchvision import transforms as T
from torchvision.transforms import functional as TF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]   # transform.py:16

# (source h, source w, output size): down- and up-scaling, mixed, identity on one axis, tiny, prime sizes
CASES = [(97, 131, 64), (131, 97, 64), (48, 64, 96), (64, 64, 64), (64, 200, 64), (7, 5, 32), (211, 89, 48), (30, 300, 96)]


def synth(rng, h, w, kind):
    if kind == 0:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127.5 + 127.5 * np.sin(xx / 7.0 + np.array([0.0, 1.0, 2.0])[:, None, None]) * np.cos(yy / 5.0)
    return np.clip(base.transpose(1, 2, 0) + rng.normal(0, 20, (h, w, 3)), 0, 255).astype(np.uint8)


def main():
    rng = np.random.default_rng(20260)
    out = {"versions": np.array([PIL.__version__, torchvision.__version__])}
    for i, (h, w, s) in enumerate(CASES):
        img = synth(rng, h, w, i % 2)
        pil = Image.fromarray(img)
        out["src_%d" % i] = img
        out["resized_%d" % i] = np.asarray(pil.resize((s, s), Image.BICUBIC))
        tr = T.Compose([T.Resize((s, s), interpolation=T.InterpolationMode.BICUBIC), T.ToTensor(), T.Normalize(MEAN, STD)])
        out["albef_%d" % i] = tr(pil).numpy()
        # albef_randaug geometry: crop + flip with torchvision's own random draws, recorded
        torch.manual_seed(100 + i)
        top, left, bh, bw = T.RandomResizedCrop.get_params(pil, [0.5, 1.0], [3 / 4, 4 / 3])
        flip = bool(torch.rand(1) < 0.5)
        c = TF.resized_crop(pil, top, left, bh, bw, (s, s), T.InterpolationMode.BICUBIC)
        if flip:
            c = TF.hflip(c)
        out["box_%d" % i] = np.array([left, top, bw, bh, int(flip)])
        out["crop_%d" % i] = T.Normalize(MEAN, STD)(T.ToTensor()(c)).numpy()
    out["cases"] = np.array(CASES)
    path = os.path.join(ROOT, "tests", "golden", "image_pipeline.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())
