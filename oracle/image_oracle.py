"""CPU oracle of the image side of FIBER's input pipeline — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; nothing under
fiber_b200/ does (tests/test_host_cpu.py enforces it).

What the reference does per image (coarse_grained/fiber/datasets/base_dataset.py:93-110 `get_raw_image` /
`get_image`, transforms/transform.py:10-45):

    PIL RGB image  ->  [RandomResizedCrop(size, scale=(0.5, 1)) + RandomHorizontalFlip (+ RandomAugment)]   albef_randaug
                   ->  Resize((size, size), BICUBIC)                                                         albef
                   ->  ToTensor()  ->  Normalize(mean, std)          (transform.py:13-17, 42-44)

The arithmetic lives in two un-vendored dependencies of the reference (coarse_grained/requirements.txt: Pillow,
torchvision), restated here from their published algorithms:

  * Pillow `Image.resize(..., BICUBIC)` on 8-bit images (libImaging/Resample.c: precompute_coeffs,
    normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc): a separable two-pass
    convolution, horizontal first, each pass rounding to uint8.  Coefficients are evaluated in double precision
    (Keys bicubic, a = -0.5, support 2 x max(scale, 1)), normalised to sum 1, converted to 22-bit fixed point with
    round-half-away-from-zero; a pass accumulates `1 << 21` + sum(pixel * coefficient) in int32, shifts right by 22
    and clamps to [0, 255].  The horizontal pass is skipped when the width does not change, the vertical one when
    the height does not (identity coefficients give the same bytes, so the restatement always runs both).
  * torchvision `ToTensor` (uint8 HWC -> float32 CHW, true division by 255) and `Normalize`
    (`sub_(mean).div_(std)` with float32 mean / std), both IEEE float32, one rounding per operation.

Pinned (tests/test_image_pipeline_cpu.py) against Pillow 12.2 and torchvision 0.26 executed in the build container —
`tools/make_golden_images.py` writes tests/golden/image_pipeline.npz from those libraries — so parity of this row
is anchored on the libraries' own outputs, bit for bit.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
MEAN = (0.485, 0.456, 0.406)      # transform.py:16
STD = (0.229, 0.224, 0.225)


def bicubic_filter(x):
    """Keys cubic convolution kernel with a = -0.5 (Resample.c: bicubic_filter)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box (in0 = 0, in1 = in_size).

    Returns (ksize, bounds int32 [out, 2] = (first tap, tap count), kk int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            if v < 0:
                kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS))
            else:
                kk[xx, x] = int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img, out_size, axis):
    """One 8-bit resampling pass along `axis` of an [H, W, C] uint8 array."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)           # [in, other, C]
    _, bounds, kk = precompute_coeffs(src.shape[0], out_size)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        lo, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n].astype(np.int64), src[lo:lo + n], axes=(0, 0))
        # int32 accumulator in the library: |sum| <= 255 * sum|k| < 2^31 for the bicubic kernel, so int64 agrees
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img, out_h, out_w, box=None, flip=False):
    """`PIL.Image.resize((out_w, out_h), BICUBIC)` of an [H, W, 3] uint8 array, after an optional crop
    box = (left, top, width, height) (torchvision `resized_crop`: crop, then resize the crop) and followed by an
    optional horizontal flip (RandomHorizontalFlip, transform.py:24)."""
    img = np.asarray(img, np.uint8)
    if box is not None:
        x0, y0, bw, bh = box
        img = img[y0:y0 + bh, x0:x0 + bw]
    tmp = _pass(img, out_w, 1)        # horizontal first (Resample.c: ImagingResampleInner)
    out = _pass(tmp, out_h, 0)
    if flip:
        out = out[:, ::-1]
    return np.ascontiguousarray(out)


def normalize_lut(mean=MEAN, std=STD):
    """float32 [3, 256]: ToTensor (v / 255) then Normalize ((x - mean) / std), one float32 rounding per operation."""
    v = np.arange(256, dtype=np.float32) / np.float32(255.0)
    m = np.asarray(mean, np.float32)[:, None]
    s = np.asarray(std, np.float32)[:, None]
    return ((v[None, :] - m) / s).astype(np.float32)


def albef_transform(img, size=384, box=None, flip=False, mean=MEAN, std=STD):
    """transform.py:10-17 on one [H, W, 3] uint8 image -> float32 [3, size, size]."""
    r = resize_bicubic_u8(img, size, size, box, flip)
    lut = normalize_lut(mean, std)
    return np.stack([lut[c][r[:, :, c]] for c in range(3)], 0)


def albef_transform_hw(img, out_h, out_w, box=None, flip=False, mean=MEAN, std=STD):
    """albef_transform with a rectangular output (Resize((out_h, out_w)))."""
    r = resize_bicubic_u8(img, out_h, out_w, box, flip)
    lut = normalize_lut(mean, std)
    return np.stack([lut[c][r[:, :, c]] for c in range(3)], 0)
