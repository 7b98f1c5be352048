"""Image side of the input pipeline: throughput of fiber_b200.transforms.albef_transform on one GPU
against the HBM roofline, with the reference's per-image CPU transform (Pillow + torchvision, the libraries
transforms/transform.py:10-17 calls) timed beside it on the host.

    python tools/bench_image.py [--batch 64] [--src 480x640] [--size 384] [--steps 20] [--warmup 5]

Workload: `--batch` decoded RGB images of `--src` pixels (COCO-sized) -> normalised float32 [B, 3, size, size], the
`batch["image"][0]` of BASELINE configs[1].  Algorithmic bytes per image = 3 h w (source box read once) + 12 size^2
(float32 output written once).  `value` times the three kernels with the sources resident in HBM (CUDA events on the
launching stream, L2 flushed between iterations); `e2e` includes packing the host images into the pinned staging
buffer and the host->device copy."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _peak_gbs():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_copy_gbs", "hbm_gbs", "hbm_GBps", "hbm_copy_GBps"):
            if k in p:
                return float(p[k]), "MEASURED_PEAKS.json"
        for k, v in p.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v), "MEASURED_PEAKS.json:%s" % k
    except Exception:  # noqa: BLE001
        pass
    return 6551.0, "SURVEY.md §9 (measured copy bandwidth)"


def _ref_one(args):
    """One image through the libraries the reference's `albef` transform calls (worker-process body)."""
    from PIL import Image
    from torchvision import transforms as T
    a, size = args
    tr = T.Compose([T.Resize((size, size), interpolation=T.InterpolationMode.BICUBIC), T.ToTensor(),
                    T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
    return float(tr(Image.fromarray(a))[0, 0, 0])


def _measure_from_jpeg(tr, host, out, size, steps, warmup, cpu_sample):
    """Encoded JPEG bytes -> batch: nvJPEG decode on the GPU (torchvision.io, a library call) + the transform, against
    PIL decode + the per-image CPU transform on one core.  The decoders differ by rounding, so this leg is timed, not
    compared bit for bit (tests/test_image_pipeline_gpu.py::test_jpeg_decode_front_end bounds the difference)."""
    try:
        import io
        from PIL import Image
        from torchvision import transforms as T
        from fiber_b200.transforms import decode_jpegs
        yy, xx = np.mgrid[0:host[0].shape[0], 0:host[0].shape[1]]
        data = []
        for i in range(len(host)):   # photo-like content (noise would make the files incompressible)
            a = np.clip(127 + 100 * (np.sin(xx / (9.0 + i % 7)) * np.cos(yy / (5.0 + i % 5)))[..., None] * np.array([1, 0.8, 0.6])
                        + host[i].astype(np.float32) * 0.05, 0, 255).astype(np.uint8)
            buf = io.BytesIO()
            Image.fromarray(a).save(buf, format="JPEG", quality=90)
            data.append(buf.getvalue())
        bufs = [torch.frombuffer(bytearray(b), dtype=torch.uint8) for b in data]
        for _ in range(warmup):
            tr(decode_jpegs(bufs), out=out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            tr(decode_jpegs(bufs), out=out)
        torch.cuda.synchronize()
        t_gpu = (time.perf_counter() - t0) / steps
        ref = T.Compose([T.Resize((size, size), interpolation=T.InterpolationMode.BICUBIC), T.ToTensor(),
                         T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
        t0 = time.perf_counter()
        for b in data[:cpu_sample]:
            ref(Image.open(io.BytesIO(b)).convert("RGB"))
        t_cpu = (time.perf_counter() - t0) / cpu_sample
        return {"value": len(data) / t_gpu, "unit": "images/s", "jpeg_bytes_per_batch": sum(len(b) for b in data),
                "decoder": "nvJPEG via torchvision.io.decode_jpeg(device='cuda') (library)", "cpu_one_core": 1.0 / t_cpu,
                "note": "host wall clock: H2D of the encoded bytes, GPU decode to planar uint8, three transform kernels"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def _all_cores_baseline(images, size):
    """The same per-image transform on every host core, one process each (the reference's DataLoader runs
    `num_workers` such processes, config.py:91).  Spawned, not forked: the parent holds a CUDA context."""
    try:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        with mp.get_context("spawn").Pool(cores, initializer=torch.set_num_threads, initargs=(1,)) as pool:
            work = [(a, size) for a in images] * max(1, (8 * cores) // len(images))
            pool.map(_ref_one, work[:cores])          # warm the workers
            t0 = time.perf_counter()
            pool.map(_ref_one, work, chunksize=1)
            return {"value": len(work) / (time.perf_counter() - t0), "unit": "images/s", "cores": cores,
                    "sample": "%d transforms over %d worker processes" % (len(work), cores)}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def measure_image_pipeline(dev, batch=64, src=(480, 640), size=384, steps=20, warmup=5, cpu_sample=16, sweep=False,
                           all_cores_baseline=False, jpeg=False):
    from fiber_b200 import lib
    from fiber_b200.transforms import albef_transform
    h, w = src
    rng = np.random.default_rng(1234)
    host = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(batch)]
    resident = [torch.from_numpy(a).to(dev) for a in host]
    tr = albef_transform(size)
    out = torch.empty(batch, 3, size, size, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def run(images, n, with_host):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        t_host = 0.0
        for a, b in ev:
            flush.zero_()
            if with_host:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tr(images, out=out)
                torch.cuda.synchronize()
                t_host += time.perf_counter() - t0
            else:
                a.record()
                tr(images, out=out)
                b.record()
        torch.cuda.synchronize()
        return t_host / n if with_host else float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e-3

    variants = {}
    if sweep:   # every "image_variant" (rows / columns per thread), same bytes out
        for v in (0, 1, 2, 3, 7, 11, 15, 19, 23):
            lib.set_option("image_variant", v)
            run(resident, warmup, False)
            variants[str(v)] = run(resident, steps, False) * 1e3
        lib.set_option("image_variant", -1)
    run(resident, warmup, False)
    l0 = lib.launch_count()
    t_dev = run(resident, steps, False)
    launches = (lib.launch_count() - l0) // steps
    pinned = [torch.from_numpy(a).pin_memory() for a in host]   # what a DataLoader with pin_memory=True hands over
    run(pinned, warmup, True)
    t_e2e = run(pinned, steps, True)
    run(host, warmup, True)
    t_e2e_pageable = run(host, steps, True)
    alg = batch * (3 * h * w + 12 * size * size)
    peak, peak_src = _peak_gbs()
    rec = {
        "metric": "images/s through the albef transform (bicubic resize + ToTensor + Normalize)",
        "value": batch / t_dev, "unit": "images/s", "ms_per_batch": t_dev * 1e3,
        "config": {"workload": "%d decoded RGB images %dx%d -> float32 [B,3,%d,%d]" % (batch, h, w, size, size),
                   "l2": "256 MB flush write between iterations"},
        "dtype": "u8 -> int32 fixed point -> f32",
        "gpu_launches_per_batch": int(launches),
        "image_variant": lib.get_option("image_variant"),
        "e2e": {"value": batch / t_e2e, "unit": "images/s", "h2d_bytes_per_step": batch * 3 * h * w,
                "d2h_bytes_per_step": 0, "from_pageable_memory": batch / t_e2e_pageable,
                "note": "host wall clock around the call + synchronize; inputs in pinned host memory: one H2D copy per image, "
                        "three kernels (from_pageable_memory: one host memcpy per image into the pinned staging buffer, "
                        "then one H2D copy)"},
        "roofline": {"bound": "hbm", "achieved": alg / t_dev / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / t_dev / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_image": 3 * h * w + 12 * size * size,
                     "note": "all three launches of a batch timed together (coefficients, horizontal, vertical pass)"},
    }
    if variants:
        rec["variant_ms_per_batch"] = variants
    if jpeg:
        rec["from_jpeg_bytes"] = _measure_from_jpeg(tr, host, out, size, steps, warmup, cpu_sample)
    try:  # the reference's own per-image path on the host: PIL resize + torchvision ToTensor / Normalize
        from PIL import Image
        from torchvision import transforms as T
        ref = T.Compose([T.Resize((size, size), interpolation=T.InterpolationMode.BICUBIC), T.ToTensor(),
                         T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
        pil = [Image.fromarray(a) for a in host[:cpu_sample]]
        nt = torch.get_num_threads()
        torch.set_num_threads(1)
        ref(pil[0])
        t0 = time.perf_counter()
        outs = [ref(p) for p in pil]
        t_cpu = (time.perf_counter() - t0) / len(pil)
        torch.set_num_threads(nt)
        same = bool(torch.equal(torch.stack(outs), tr(host[:cpu_sample]).cpu()))
        all_cores = None
        if all_cores_baseline:
            all_cores = _all_cores_baseline(host[:cpu_sample], size)
        rec["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "images/s", "cores": 1, "kind": "reference", "all_cores": all_cores,
                               "sample": "%d of the batch's images through Pillow %s resize + torchvision ToTensor/Normalize "
                                         "(what transforms/transform.py:10-17 runs per image in a DataLoader worker)"
                                         % (len(pil), Image.__version__),
                               "gpu_output_identical": same}
    except Exception as e:  # noqa: BLE001
        rec["cpu_baseline"] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--src", default="480x640")
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--sweep", action="store_true", help="also time every image_variant")
    a = ap.parse_args()
    h, w = (int(v) for v in a.src.split("x"))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    print(json.dumps(measure_image_pipeline(dev, a.batch, (h, w), a.size, a.steps, a.warmup, sweep=a.sweep, all_cores_baseline=True, jpeg=True)))


if __name__ == "__main__":
    main()
