"""Image side of the input pipeline, CPU half:
  * the oracle (oracle/image_oracle.py) against the golden outputs of Pillow / torchvision (tests/golden/
    image_pipeline.npz, tools/make_golden_images.py) and against the live libraries when they import;
  * the kernels' per-element bodies (fiber_b200/csrc/image_resample.cuh) replayed on the CPU with the launcher's grid
    arithmetic (tests/native/image_pipeline_emul.cpp), descriptors planned by the real library, against the oracle;
  * the host logic of fiber_b200.transforms (random draws, plan errors).
No compute call into the CUDA library is made here."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import image_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "image_pipeline.npz"))
CASES = [tuple(int(v) for v in c) for c in GOLD["cases"]]


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_library_golden(i):
    h, w, s = CASES[i]
    src = GOLD["src_%d" % i]
    assert src.shape == (h, w, 3)
    assert np.array_equal(O.resize_bicubic_u8(src, s, s), GOLD["resized_%d" % i])
    assert np.array_equal(O.albef_transform(src, s), GOLD["albef_%d" % i])          # float32, bit for bit
    left, top, bw, bh, flip = (int(v) for v in GOLD["box_%d" % i])
    assert np.array_equal(O.albef_transform(src, s, box=(left, top, bw, bh), flip=bool(flip)), GOLD["crop_%d" % i])


def test_oracle_matches_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    for h, w, oh, ow in [(120, 160, 96, 96), (33, 47, 64, 32), (50, 50, 50, 80), (400, 300, 96, 128), (9, 300, 24, 24)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(O.resize_bicubic_u8(img, oh, ow), ref), (h, w, oh, ow)


def test_normalize_lut_matches_torch():
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 256, 1).expand(1, 256, 3).contiguous()   # H=1, W=256, C=3
    t = ramp.permute(2, 0, 1).to(torch.float32).div(255)                                          # ToTensor
    mean, std = torch.tensor(O.MEAN).view(3, 1, 1), torch.tensor(O.STD).view(3, 1, 1)
    ref = t.sub(mean).div(std)[:, 0, :].numpy()                                                   # Normalize
    assert np.array_equal(O.normalize_lut(), ref)


# ---- the kernels' bodies, replayed --------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    so = tmp_path_factory.mktemp("img_emul") / "libimg_emul.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "native", "image_pipeline_emul.cpp"), "-o", str(so)], check=True)
    return C.CDLL(str(so))


def _run_emul(emul, images, out_h, out_w, boxes=None, flips=None, strides=None, variant=0, planar=None):
    from fiber_b200 import lib as L
    lib = L.load()
    n = len(images)
    descs = (L.ImageDesc * n)()
    keep = []
    for i, img in enumerate(images):
        h, w = img.shape[:2]
        stride = strides[i] if strides else 3 * w
        lead = 16 + i % 4       # every byte alignment of the first pixel; padding for the word-form pass's aligned words
        d = descs[i]
        if planar and planar[i]:   # three byte planes [3][h][w + pad], a gap between the planes
            pstride, gap = w + i % 3, 5 * (i % 2)
            raw = np.full(lead + 3 * (h * pstride + gap) + 16, 0xAB, np.uint8)
            for c in range(3):
                o = lead + c * (h * pstride + gap)
                raw[o:o + h * pstride].reshape(h, pstride)[:, :w] = img[:, :, c]
            keep.append(raw)
            d.src, d.stride, d.h, d.w = raw.ctypes.data + lead, pstride, h, w
            d.planar, d.chan_stride = 1, h * pstride + gap
        else:
            raw = np.full(lead + h * stride + 16, 0xAB, np.uint8)
            buf = raw[lead:lead + h * stride].reshape(h, stride)
            buf[:, :3 * w] = img.reshape(h, 3 * w)
            keep.append(raw)
            d.src, d.stride, d.h, d.w = buf.ctypes.data, stride, h, w
        d.box_x, d.box_y, d.box_w, d.box_h = boxes[i] if boxes else (0, 0, w, h)
        d.flip = int(flips[i]) if flips else 0
    need = lib.fiber_image_transform_plan(descs, n, out_h, out_w)
    assert need > 0, lib.fiber_last_error()
    ws = np.zeros(need + 16, np.uint8)
    base = (ws.ctypes.data + 15) // 16 * 16
    out = np.full((n, 3, out_h, out_w), np.nan, np.float32)
    mean, std = (C.c_float * 3)(*O.MEAN), (C.c_float * 3)(*O.STD)
    emul.emul_image_transform(descs, n, out_h, out_w, mean, std, C.c_void_p(base), C.c_void_p(out.ctypes.data), variant)
    return out


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 7, 11, 15, 18, 19, 23])
def test_kernel_bodies_match_oracle_ragged_batch(emul, variant):
    rng = np.random.default_rng(11)
    sizes = [(97, 131), (48, 64), (64, 64), (7, 5), (211, 89), (30, 300), (64, 65), (1, 1), (130, 64)]
    images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    for out_h, out_w in [(64, 64), (96, 32), (40, 72), (24, 12)]:
        got = _run_emul(emul, images, out_h, out_w, strides=[3 * w + (5 * i) % 7 for i, (h, w) in enumerate(sizes)],
                        variant=variant)
        for i, img in enumerate(images):
            r = O.resize_bicubic_u8(img, out_h, out_w)
            lut = O.normalize_lut()
            want = np.stack([lut[c][r[:, :, c]] for c in range(3)], 0)
            assert np.array_equal(got[i], want), (sizes[i], out_h, out_w)


@pytest.mark.parametrize("variant", [0, 3, 19])
def test_kernel_bodies_match_library_golden_with_crop_and_flip(emul, variant):
    for i, (h, w, s) in enumerate(CASES):
        src = GOLD["src_%d" % i]
        left, top, bw, bh, flip = (int(v) for v in GOLD["box_%d" % i])
        got = _run_emul(emul, [src, src], s, s, boxes=[(0, 0, w, h), (left, top, bw, bh)], flips=[0, flip],
                        variant=variant)
        assert np.array_equal(got[0], GOLD["albef_%d" % i])
        assert np.array_equal(got[1], GOLD["crop_%d" % i])


@pytest.mark.parametrize("variant", [0, 2, 18])
def test_kernel_bodies_planar_sources(emul, variant):
    """Three byte planes per image (what GPU JPEG decoders return) next to interleaved ones in the same batch."""
    rng = np.random.default_rng(12)
    sizes = [(97, 131), (48, 64), (7, 5), (130, 64), (33, 200)]
    images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    boxes = [(3, 2, w - 5, h - 4) if min(h, w) > 10 else (0, 0, w, h) for h, w in sizes]
    got = _run_emul(emul, images, 40, 72, boxes=boxes, flips=[0, 1, 0, 1, 1], variant=variant,
                    planar=[True, False, True, True, True])
    for i, img in enumerate(images):
        assert np.array_equal(got[i], O.albef_transform_hw(img, 40, 72, box=boxes[i], flip=bool([0, 1, 0, 1, 1][i]))), sizes[i]


def test_plan_rejects_bad_descriptors():
    from fiber_b200 import lib as L
    lib = L.load()
    buf = np.zeros((10, 30), np.uint8)
    d = (L.ImageDesc * 1)()
    d[0].src, d[0].stride, d[0].h, d[0].w = buf.ctypes.data, 30, 10, 10
    d[0].box_x, d[0].box_y, d[0].box_w, d[0].box_h = 0, 0, 10, 10
    assert lib.fiber_image_transform_plan(d, 1, 8, 8) > 0
    assert d[0].ksize_x == 2 * 3 + 1 and d[0].ksize_y == 7           # scale 1.25: ceil(2.5) * 2 + 1
    assert lib.fiber_image_transform_plan(d, 1, 8, 6) == 0           # out_w % 4
    d[0].box_w = 11
    assert lib.fiber_image_transform_plan(d, 1, 8, 8) == 0           # box outside the image
    assert b"crop box" in lib.fiber_last_error()
    d[0].box_w, d[0].stride = 10, 29
    assert lib.fiber_image_transform_plan(d, 1, 8, 8) == 0           # stride shorter than a row
    tall = (L.ImageDesc * 1)()
    tall[0].src, tall[0].stride, tall[0].h, tall[0].w = buf.ctypes.data, 3, 1000, 1
    tall[0].box_w, tall[0].box_h = 1, 1000
    assert lib.fiber_image_transform_plan(tall, 1, 8, 8) == 0        # Pillow's rows-first special case
    assert lib.fiber_image_transform_plan(d, 0, 8, 8) == 0
    d[0].stride, d[0].planar, d[0].chan_stride = 10, 1, 99           # planes overlap: chan_stride < (h - 1) stride + w
    assert lib.fiber_image_transform_plan(d, 1, 8, 8) == 0
    d[0].chan_stride = 100
    assert lib.fiber_image_transform_plan(d, 1, 8, 8) > 0


def test_random_draws_follow_torchvision_modules():
    T = pytest.importorskip("torchvision.transforms")
    Image = pytest.importorskip("PIL.Image")
    from fiber_b200.transforms import albef_transform_randaug, keys_to_transforms
    tr = albef_transform_randaug(96)
    assert [type(t).__name__ for t in keys_to_transforms(["albef", "albef_randaug"], size=64)] == ["BatchImageTransform"] * 2
    sizes = [(120, 160), (90, 60), (300, 200)]
    torch.manual_seed(7)
    boxes, flips = tr.draw_params(sizes)
    # the same seed through torchvision's own modules, image by image (transform.py:22-24)
    torch.manual_seed(7)
    rng = np.random.default_rng(0)
    crop, flipm = T.RandomResizedCrop(96, scale=(0.5, 1.0), interpolation=T.InterpolationMode.BICUBIC), T.RandomHorizontalFlip()
    for (h, w), box, flip in zip(sizes, boxes, flips):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(flipm(crop(Image.fromarray(img))))
        assert np.array_equal(O.resize_bicubic_u8(img, 96, 96, box=box, flip=flip), ref)


def test_transform_needs_cuda():
    from fiber_b200.transforms import albef_transform
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        albef_transform(32)([np.zeros((8, 8, 3), np.uint8)])


def test_image_layout_detection_on_the_host():
    """fiber_b200.transforms._as_image: what counts as interleaved / planar, and what is rejected (host logic only)."""
    from fiber_b200.transforms import _as_image
    hwc = np.zeros((5, 7, 3), np.uint8)
    t, planar = _as_image(hwc)
    assert not planar and tuple(t.shape) == (5, 7, 3) and t.stride(2) == 1 and t.stride(1) == 3
    chw = torch.zeros(3, 5, 7, dtype=torch.uint8)
    t, planar = _as_image(chw)
    assert planar and tuple(t.shape) == (3, 5, 7) and t.stride(2) == 1
    wide = torch.zeros(5, 9, 3, dtype=torch.uint8)[:, :7]            # row stride > 3 w: read in place
    t, planar = _as_image(wide)
    assert not planar and t.data_ptr() == wide.data_ptr() and t.stride(0) == 27
    sliced = torch.zeros(3, 5, 14, dtype=torch.uint8)[:, :, ::2]     # pixel stride 2: copied
    t, planar = _as_image(sliced)
    assert planar and t.stride(2) == 1
    Image = pytest.importorskip("PIL.Image")
    t, planar = _as_image(Image.fromarray(np.zeros((4, 6), np.uint8)))   # grey PIL image -> RGB, as get_raw_image does
    assert not planar and tuple(t.shape) == (4, 6, 3)
    for bad in (np.zeros((5, 7, 3), np.float32), np.zeros((5, 7), np.uint8), np.zeros((5, 7, 4), np.uint8)):
        with pytest.raises(RuntimeError, match="uint8"):
            _as_image(bad)


def test_transform_keys_mirror_the_reference():
    """Same keys as coarse_grained/fiber/transforms/__init__.py:6-9 (parsed from the reference when it is present)."""
    from fiber_b200 import transforms as T
    assert set(T._transforms) == {"albef", "albef_randaug"}
    ref = "/root/reference/coarse_grained/fiber/transforms/__init__.py"
    if os.path.exists(ref):
        import re
        keys = set(re.findall(r'"(\w+)":\s*\w+', open(ref).read()))
        assert keys == set(T._transforms)
