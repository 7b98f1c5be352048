"""Image side of the input pipeline on the GPU, through fiber_b200.transforms -> the C-ABI
(fiber_image_transform): bit-exact against the oracle and against the golden outputs of Pillow / torchvision."""
import os

import numpy as np
import pytest
import torch

from oracle import image_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "image_pipeline.npz"))
CASES = [tuple(int(v) for v in c) for c in GOLD["cases"]]


def _launches():
    from fiber_b200 import lib
    return lib.launch_count()


@pytest.mark.parametrize("on_device", [False, True])
def test_golden_cases(on_device):
    from fiber_b200.transforms import BatchImageTransform
    for s in sorted({c[2] for c in CASES}):
        idx = [i for i, c in enumerate(CASES) if c[2] == s]
        srcs = [GOLD["src_%d" % i] for i in idx]
        boxes = [tuple(int(v) for v in GOLD["box_%d" % i][:4]) for i in idx]
        flips = [int(GOLD["box_%d" % i][4]) for i in idx]
        imgs = [torch.from_numpy(a).cuda() for a in srcs] if on_device else srcs
        tr = BatchImageTransform(s)
        before = _launches()
        full = tr(imgs).cpu().numpy()
        assert _launches() - before == 3
        crop = tr(imgs, boxes=boxes, flips=flips).cpu().numpy()
        for j, i in enumerate(idx):
            assert np.array_equal(full[j], GOLD["albef_%d" % i]), ("albef", CASES[i])
            assert np.array_equal(crop[j], GOLD["crop_%d" % i]), ("crop", CASES[i])


@pytest.fixture
def variant(request):
    from fiber_b200 import lib
    lib.set_option("image_variant", request.param)
    yield request.param
    lib.set_option("image_variant", -1)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 7, 11, 15, 19, 23], indirect=True)
def test_ragged_batch_vs_oracle_rectangular_output_and_strides(variant):
    from fiber_b200.transforms import BatchImageTransform
    rng = np.random.default_rng(3)
    sizes = [(97, 131), (48, 64), (64, 64), (7, 5), (211, 89), (30, 300), (64, 65), (1, 1), (130, 64), (500, 375)]
    images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    # device images as views into wider buffers: row stride > 3 w
    wide = [torch.from_numpy(np.pad(a, ((0, 0), (0, 3 + i), (0, 0)))).cuda()[:, :a.shape[1]] for i, a in enumerate(images)]
    for out_hw in [(64, 64), (96, 32), (40, 72), (24, 12)]:
        tr = BatchImageTransform(out_hw)
        for batch in (images, wide):
            got = tr(batch).cpu().numpy()
            for i, img in enumerate(images):
                assert np.array_equal(got[i], O.albef_transform_hw(img, *out_hw)), (sizes[i], out_hw)


def test_random_crop_same_seed_as_torchvision():
    T = pytest.importorskip("torchvision.transforms")
    Image = pytest.importorskip("PIL.Image")
    from fiber_b200.transforms import albef_transform_randaug
    rng = np.random.default_rng(8)
    images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(120, 160), (90, 60), (300, 200), (64, 64)]]
    torch.manual_seed(21)
    got = albef_transform_randaug(96)(images).cpu().numpy()
    torch.manual_seed(21)
    ref_tr = T.Compose([T.RandomResizedCrop(96, scale=(0.5, 1.0), interpolation=T.InterpolationMode.BICUBIC),
                        T.RandomHorizontalFlip(), T.ToTensor(), T.Normalize(O.MEAN, O.STD)])   # transform.py:20-44 minus RandomAugment
    for i, img in enumerate(images):
        assert np.array_equal(got[i], ref_tr(Image.fromarray(img)).numpy()), i


def test_full_size_batch_properties_and_buffer_reuse():
    """BASELINE configs[1] input shape: 64 images -> 384 x 384.  Spot-checks against the oracle plus size-independent
    properties: an image already at the output size passes through the look-up table untouched, a constant image stays
    constant, and a second call that reuses the staging buffers gives the same bytes."""
    from fiber_b200.transforms import albef_transform
    rng = np.random.default_rng(1)
    B, S = 64, 384
    sizes = [(480, 640), (640, 480), (384, 384), (333, 500), (512, 512), (768, 1024), (240, 320), (427, 640)]
    images = []
    for i in range(B):
        h, w = sizes[i % len(sizes)]
        images.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if i != 5 else np.full((h, w, 3), (17, 130, 255), np.uint8))
    tr = albef_transform(S)
    out = torch.empty(B, 3, S, S, device="cuda")
    got = tr(images, out=out)
    assert got.data_ptr() == out.data_ptr()
    got = got.cpu().numpy()
    lut = O.normalize_lut()
    for i in (0, 1, 3, 7, 63):
        assert np.array_equal(got[i], O.albef_transform(images[i], S)), i
    for i in range(2, B, len(sizes)):                      # 384 x 384 sources: identity resampling
        want = np.stack([lut[c][images[i][:, :, c]] for c in range(3)], 0)
        assert np.array_equal(got[i], want), i
    for c, v in enumerate((17, 130, 255)):
        assert np.all(got[5][c] == lut[c][v])
    again = tr(list(reversed(images))).cpu().numpy()
    assert np.array_equal(again[::-1], got)


def test_pinned_pageable_and_device_images_in_one_batch():
    from fiber_b200.transforms import BatchImageTransform
    rng = np.random.default_rng(4)
    images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(90, 120), (64, 48), (33, 77), (120, 90), (50, 50), (71, 19)]]
    mixed = [torch.from_numpy(images[0]).pin_memory(), images[1], torch.from_numpy(images[2]).cuda(),
             torch.from_numpy(images[3]).pin_memory(), torch.from_numpy(images[4]), torch.from_numpy(images[5]).pin_memory()]
    tr = BatchImageTransform(64)
    for _ in range(2):   # second call reuses the staging buffers
        got = tr(mixed).cpu().numpy()
        for i, img in enumerate(images):
            assert np.array_equal(got[i], O.albef_transform(img, 64)), i


def test_planar_sources_equal_interleaved():
    from fiber_b200.transforms import BatchImageTransform
    rng = np.random.default_rng(6)
    images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(90, 120), (64, 48), (333, 500), (71, 19)]]
    chw = [torch.from_numpy(a).permute(2, 0, 1).contiguous() for a in images]
    mixed = [chw[0].cuda(), images[1], chw[2], chw[3].pin_memory()]
    tr = BatchImageTransform((64, 96))
    got = tr(mixed).cpu().numpy()
    for i, img in enumerate(images):
        assert np.array_equal(got[i], O.albef_transform_hw(img, 64, 96)), i
    torch.manual_seed(3)
    a = BatchImageTransform(64, random_crop=True)(images).cpu()
    torch.manual_seed(3)
    b = BatchImageTransform(64, random_crop=True)([c.cuda() for c in chw]).cpu()
    assert torch.equal(a, b)


def test_jpeg_decode_front_end():
    """nvJPEG through torchvision, then the exact transform: equals the transform of the SAME decoded pixels bit for bit,
    and stays within decoder rounding of the PIL-decoded path."""
    Image = pytest.importorskip("PIL.Image")
    import io
    from fiber_b200.transforms import albef_transform, decode_jpegs
    rng = np.random.default_rng(9)
    yy, xx = np.mgrid[0:240, 0:320]
    imgs = [np.clip(127 + 100 * (np.sin(xx / (9.0 + i)) * np.cos(yy / 7.0))[..., None] * np.array([1, 0.8, 0.6])
                    + rng.normal(0, 4, (240, 320, 3)), 0, 255).astype(np.uint8) for i in range(3)]
    data = []
    for a in imgs:
        buf = io.BytesIO()
        Image.fromarray(a).save(buf, format="JPEG", quality=92, subsampling=0)
        data.append(buf.getvalue())
    try:
        decoded = decode_jpegs(data)
    except Exception as e:  # noqa: BLE001  (torchvision built without nvJPEG)
        pytest.skip("GPU JPEG decode unavailable: %s" % e)
    assert all(t.is_cuda and t.dtype == torch.uint8 and tuple(t.shape) == (3, 240, 320) for t in decoded)
    tr = albef_transform(96)
    got = tr(decoded).cpu().numpy()
    for i, t in enumerate(decoded):
        assert np.array_equal(got[i], O.albef_transform(t.permute(1, 2, 0).cpu().numpy(), 96)), i
    pil = [np.asarray(Image.open(io.BytesIO(b)).convert("RGB")) for b in data]
    ref = tr(pil).cpu().numpy()
    lut = O.normalize_lut()
    step = float(lut[0][1] - lut[0][0])                       # one grey level after normalisation
    diff = np.abs(got - ref)
    print("nvJPEG vs libjpeg after the transform: max %.2f grey levels, mean %.3f" % (diff.max() / step, diff.mean() / step))
    assert diff.max() <= 8.5 * step and diff.mean() <= 1.0 * step


def test_errors_are_loud():
    from fiber_b200.transforms import BatchImageTransform
    with pytest.raises(RuntimeError, match="uint8"):
        BatchImageTransform(32)([np.zeros((8, 8, 3), np.float32)])
    with pytest.raises(RuntimeError, match="image_transform_plan"):
        BatchImageTransform(32)([np.zeros((8, 8, 3), np.uint8)], boxes=[(4, 4, 8, 8)], flips=[0])
    with pytest.raises(RuntimeError, match="image_transform_plan"):
        BatchImageTransform((32, 30))([np.zeros((8, 8, 3), np.uint8)])
