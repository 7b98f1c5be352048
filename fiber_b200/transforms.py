"""Image side of the input pipeline on the GPU — host mirror of `fiber.transforms`
(coarse_grained/fiber/transforms/__init__.py:1-13, transform.py:10-45).

The reference transforms ONE PIL image at a time on DataLoader worker cores (datasets/base_dataset.py:93-110):
`albef` = Resize((size, size), BICUBIC) -> ToTensor -> Normalize, `albef_randaug` = RandomResizedCrop(size,
scale=(0.5, 1), BICUBIC) -> RandomHorizontalFlip -> RandomAugment -> ToTensor -> Normalize.  Here the same names
return a BATCH transform: a list of decoded uint8 RGB images of any sizes (interleaved [H, W, 3] or planar [3, H, W]) goes in, the normalised float32
`[B, 3, size, size]` CUDA tensor that `batch["image"][0]` holds comes out of three kernel launches
(csrc/image_pipeline.cu), bit-identical to the reference's per-image result (Pillow's fixed-point bicubic
resampling and torchvision's float32 normalisation are reproduced exactly; tests/test_image_pipeline_*.py).

Scope: decoding (JPEG -> uint8 RGB) stays with the caller — `decode_jpegs` offers nvJPEG through torchvision for it, whose
pixels differ from libjpeg's by decoder rounding; `albef_randaug` here applies the crop and the flip with
the same torch RNG draws, in the same order, as torchvision's modules — RandomAugment's ten photometric / affine
operations (transforms/randaug.py) are not applied (`BatchImageTransform.randaug_ops` is False and documents it).
There is no CPU path: without the CUDA library every call raises.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _lib

MEAN = (0.485, 0.456, 0.406)  # transform.py:16
STD = (0.229, 0.224, 0.225)


def _as_image(img):
    """PIL image / ndarray / tensor -> (uint8 tensor, planar?): [H, W, 3] with unit channel stride, or three byte planes
    [3, H, W] with unit pixel stride (what GPU JPEG decoders return); host or CUDA."""
    if isinstance(img, torch.Tensor):
        t = img
    elif isinstance(img, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(img))
    else:  # PIL.Image (get_raw_image returns .convert("RGB"), base_dataset.py:99)
        t = torch.from_numpy(np.asarray(img.convert("RGB")).copy())
    if t.dtype != torch.uint8 or t.dim() != 3 or (t.shape[2] != 3 and t.shape[0] != 3):
        raise RuntimeError("fiber_b200.transforms: images must be uint8 [H, W, 3] or [3, H, W] (decoded RGB)")
    planar = t.shape[2] != 3
    if planar:
        if t.stride(2) != 1 or t.stride(0) < (t.shape[1] - 1) * t.stride(1) + t.shape[2]:
            t = t.contiguous()
    elif t.stride(2) != 1 or t.stride(1) != 3:
        t = t.contiguous()
    return t, planar


def decode_jpegs(data, device=None):
    """Encoded JPEG byte strings -> list of uint8 [3, H, W] CUDA tensors, decoded on the GPU by nvJPEG through
    torchvision.io (a LIBRARY call, the counterpart of `Image.open(...).convert("RGB")`, base_dataset.py:93-99).
    nvJPEG's inverse DCT and chroma upsampling are not bit-identical to libjpeg's, so a batch built from these pixels
    differs from the reference's by decoder rounding (a few grey levels on few pixels); everything after the decode is exact."""
    from torchvision.io import ImageReadMode, decode_jpeg
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    bufs = [torch.frombuffer(bytearray(b), dtype=torch.uint8) if isinstance(b, (bytes, bytearray, memoryview)) else b
            for b in data]
    return decode_jpeg(bufs, mode=ImageReadMode.RGB, device=dev)


class BatchImageTransform:
    """Callable: list of decoded images -> float32 [B, 3, size, size] on the current CUDA device."""

    randaug_ops = False  # RandomAugment (transforms/randaug.py) is not part of this transform

    def __init__(self, size=384, random_crop=False, mean=MEAN, std=STD, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
                 flip_p=0.5):
        self.size = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
        self.random_crop = bool(random_crop)
        self.scale, self.ratio, self.flip_p = tuple(scale), tuple(ratio), float(flip_p)
        self._mean = (C.c_float * 3)(*mean)
        self._std = (C.c_float * 3)(*std)
        self._ws = None       # device workspace, grown on demand
        self._stage = None    # pinned host staging for host images + descriptors
        self._stage_dev = None
        self._stage_evt = None  # the previous call's host->device copy of the staging buffer

    # -- random parameters, drawn exactly as torchvision's modules draw them (transform.py:22-24) -------------------
    def draw_params(self, sizes):
        """[(h, w)] -> ([(left, top, width, height)], [flip]) consuming torch's global RNG like
        RandomResizedCrop.forward followed by RandomHorizontalFlip.forward, image by image."""
        from torchvision.transforms import RandomResizedCrop
        boxes, flips = [], []
        for h, w in sizes:
            top, left, bh, bw = RandomResizedCrop.get_params(torch.empty(3, h, w, device="meta"), list(self.scale),
                                                             list(self.ratio))
            boxes.append((left, top, bw, bh))
            flips.append(bool(torch.rand(1) < self.flip_p))
        return boxes, flips

    def _grow(self, name, nbytes, **kw):
        buf = getattr(self, name)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, **kw)
            setattr(self, name, buf)
        return buf

    def __call__(self, images, boxes=None, flips=None, out=None):
        if not torch.cuda.is_available():
            raise RuntimeError("fiber_b200.transforms: CUDA device required (there is no CPU path)")
        lib = _lib.load()
        pairs = [_as_image(i) for i in images]
        imgs, planar = [p[0] for p in pairs], [p[1] for p in pairs]
        n = len(imgs)
        if n == 0:
            raise RuntimeError("fiber_b200.transforms: empty batch")
        if self.random_crop and boxes is None:
            boxes, flips = self.draw_params([(int(t.shape[1]), int(t.shape[2])) if pl else (int(t.shape[0]), int(t.shape[1]))
                                             for t, pl in zip(imgs, planar)])
        dev = torch.device("cuda", torch.cuda.current_device())
        oh, ow = self.size
        # host images share one pinned staging buffer and one host->device copy; device images are read in place
        dsz = C.sizeof(_lib.ImageDesc)
        offs, total = [], (n * dsz + 255) // 256 * 256
        for t in imgs:
            if t.is_cuda:
                offs.append(None)
            else:
                offs.append(total)
                total += (t.numel() + 255) // 256 * 256
        if self._stage_evt is not None:
            self._stage_evt.synchronize()  # the staging buffer is about to be overwritten on the host
        stage = self._grow("_stage", total, pin_memory=True)
        stage_dev = self._grow("_stage_dev", total, device=dev)
        stage_np = stage.numpy()
        descs = (_lib.ImageDesc * n).from_buffer(stage_np)   # descriptors live at the head of the staging buffer
        keep, direct = [], []
        for i, t in enumerate(imgs):
            pl = planar[i]
            h, w = (int(t.shape[1]), int(t.shape[2])) if pl else (int(t.shape[0]), int(t.shape[1]))
            d = descs[i]
            d.planar = int(pl)
            if offs[i] is None:
                if t.device != dev:
                    t = t.to(dev)
                keep.append(t)
                d.src = t.data_ptr()
                d.stride, d.chan_stride = (t.stride(1), t.stride(0)) if pl else (t.stride(0), 0)
            else:
                if t.is_pinned() and t.is_contiguous():   # already page-locked (DataLoader pin_memory): no host memcpy
                    direct.append((offs[i], t))
                else:
                    stage[offs[i]:offs[i] + t.numel()].view(t.shape).copy_(t)
                d.src = stage_dev.data_ptr() + offs[i]
                d.stride, d.chan_stride = (w, h * w) if pl else (3 * w, 0)
            d.h, d.w = h, w
            d.box_x, d.box_y, d.box_w, d.box_h = boxes[i] if boxes is not None else (0, 0, w, h)
            d.flip = int(bool(flips[i])) if flips is not None else 0
        need = lib.fiber_image_transform_plan(descs, n, oh, ow)
        if need == 0:
            raise RuntimeError("fiber_b200.image_transform_plan failed: %s" % lib.fiber_last_error().decode())
        if direct:   # descriptors + the pageable images' staging area in one copy, pinned images one copy each
            head = max([n * dsz] + [o + imgs[i].numel() for i, o in enumerate(offs)
                                    if o is not None and not (imgs[i].is_pinned() and imgs[i].is_contiguous())])
            stage_dev[:head].copy_(stage[:head], non_blocking=True)
            for o, t in direct:
                stage_dev[o:o + t.numel()].view(t.shape).copy_(t, non_blocking=True)
        else:
            stage_dev[:total].copy_(stage[:total], non_blocking=True)
        self._stage_evt = torch.cuda.Event()
        self._stage_evt.record()
        ws = self._grow("_ws", need, device=dev)
        if out is None:
            out = torch.empty(n, 3, oh, ow, dtype=torch.float32, device=dev)
        elif out.shape != (n, 3, oh, ow) or out.dtype != torch.float32 or not out.is_contiguous() or not out.is_cuda:
            raise RuntimeError("fiber_b200.transforms: out must be a contiguous float32 CUDA tensor [B, 3, H, W]")
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.fiber_image_transform(descs, C.c_void_p(stage_dev.data_ptr()), n, oh, ow, self._mean, self._std,
                                             C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(out.data_ptr()), stream),
                   "image_transform")
        del keep
        return out


def albef_transform(size=384):
    """transform.py:10-17 as a batch transform."""
    return BatchImageTransform(size, random_crop=False)


def albef_transform_randaug(size=384):
    """transform.py:20-45 as a batch transform: crop + flip + resize + normalise (see the module docstring)."""
    return BatchImageTransform(size, random_crop=True)


_transforms = {"albef": albef_transform, "albef_randaug": albef_transform_randaug}


def keys_to_transforms(keys, size=384):
    """transforms/__init__.py:12-13."""
    return [_transforms[key](size=size) for key in keys]
