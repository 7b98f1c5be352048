#!/usr/bin/env python
"""bench.py — FIBER-Base fwd+bwd throughput (image-text pairs/s) on B200, BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): coarse pre-training step ITM+ITC+MLM, 384 px, 40 tokens,
B=64 pairs per GPU, bf16 activations / fp32 accumulation, synthetic data, random-init weights.
One step = FIBERTransformerSS.training_step (4 fused + 1 image-only + 1 text-only backbone passes
+ heads) followed by backward; the optimizer step is not part of the metric; with N>1 the gradient
all-reduce (DDP over NCCL) is inside the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import warnings

import torch

# synthetic benchmark: random-init weights are the point (`data: synthetic`), not an accident
warnings.filterwarnings("ignore", message="fiber_b200: RobertaModel.from_pretrained", category=RuntimeWarning)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_PAIR = 1636.23e9  # fwd+bwd, config 1 (FlopCounterMode over the reference)
METRIC = "image-text pairs/sec/GPU FIBER-Base 384px fwd+bwd at 1/2/4/8 B200"
WORKLOAD = "FIBER-Base coarse pretrain step ITM+ITC+MLM (BASELINE configs[1]) 384px/40tok fwd+bwd"


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from `ncu --set full`
# (profiles/r1_gemm_plain_ncu.txt): the step's largest-FLOP shape, Swin stage-2 fc2 dgrad over the 4B-sample pass.
GEMM_NCU_TRAFFIC = {"dram_bytes": 740.4e6,
                    "note": "ncu capture of one launch M=147456 N=512 K=2048 (bf16 out): 606.2 MB read + 134.3 MB written "
                            "vs 757 MB algorithmic (A 604 + B 2 + C 151); part of C was still in L2 at kernel end"}


def config(tasks, image_size=384, max_text_len=40):
    """coarse_grained/fiber/config.py:21-92 defaults + task_pretrain_mlm_itm_itc (:95-110)."""
    loss_names = {"itm": 0, "mlm": 0, "itc": 0, "vqa": 0, "nlvr2": 0, "caption_mle": 0, "caption_gold": 0,
                  "caption_cider": 0}
    loss_names.update({t: 1 for t in tasks})
    return dict(loss_names=loss_names, image_size=image_size, vit="swin_base_patch4_window12_384_in22k",
                input_image_embed_size=1024, input_text_embed_size=768, pretrained_vit=False, vqav2_label_size=3129,
                max_text_len=max_text_len, tokenizer="roberta-base", vocab_size=50265, hidden_size=768, num_heads=12,
                num_layers=12, mlp_ratio=4, drop_rate=0.1, num_fuse_block=6, itc_pooler=True, load_path="",
                test_only=False, optim_type="adamw", learning_rate=1e-5, weight_decay=0.01, decay_power=1,
                max_steps=100000, warmup_steps=10000, end_lr=0, lr_mult_head=5, lr_mult_cross_modal=5)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active")
                                                         for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def make_batch(B, image_size, L, seed=1234, vocab=50265):
    """Synthetic batch with the schema of the reference's collate (datasets/base_dataset.py:172-245),
    SURVEY.md §8(d): N(0,1) images; <s> ... </s> token ids with every other row padded; 15 % <mask>."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    batch = {"image": [torch.randn(B, 3, image_size, image_size, generator=g)]}
    ids = torch.randint(3, vocab - 1, (B, L), generator=g)
    masks = torch.ones(B, L, dtype=torch.long)
    lens = torch.randint(8, L + 1, (B,), generator=g)
    for b in range(B):
        n = int(lens[b]) if b % 2 == 1 else L
        ids[b, 0], ids[b, n - 1] = 0, 2
        ids[b, n:] = 1
        masks[b, n:] = 0
    pick = (torch.rand(B, L, generator=g) < 0.15) & (ids > 2)
    pick[:, 1] = True
    batch.update(text_ids=ids, text_masks=masks, text_labels=torch.full((B, L), -100),
                 text_ids_mlm=torch.where(pick, torch.full_like(ids, vocab - 1), ids),
                 text_labels_mlm=torch.where(pick, ids, torch.full_like(ids, -100)), text=["synthetic caption"] * B)
    return batch


# Other BASELINE.json configs, measured as EXTRA keys of the N = 1 line (never the headline): name -> (tasks, image
# size, text length, per-GPU batch, fwd+bwd GFLOP per pair from SURVEY.md §8a)
EXTRA_CONFIGS = {
    "vqa576": (["vqa"], 576, 50, 32, 771.02e9),   # configs[2]: VQAv2 fine-tuning, 18 x 18 = 324-token windows
    "itc384": (["itc"], 384, 50, 64, 309.02e9),   # configs[3]: ITC-only retrieval fine-tuning, 64 x (64 + 4096) sims
}


def measure_extra_config(name, dev, steps=3, warmup=3):
    """fwd+bwd pairs/s of one extra config on this GPU (device-resident synthetic batch, CUDA events)."""
    from fiber_b200 import ops
    from fiber_b200.modules import FIBERTransformerSS, fiber_utils
    tasks, R, L, B, flops = EXTRA_CONFIGS[name]
    torch.manual_seed(1234)
    model = FIBERTransformerSS(config(tasks, R, L)).to(dev)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(("alpha_i2t", "alpha_t2i")):
                p.fill_(0.5)
    fill_queues(model)
    model.train()
    fiber_utils.set_task(model)
    ops.set_dropout_seed(1234)
    batch = make_batch(B, R, L, seed=1234)
    if "vqa" in tasks:
        g = torch.Generator(device="cpu").manual_seed(99)
        batch["vqa_labels"] = [[int(torch.randint(0, 3129, (1,), generator=g))] for _ in range(B)]
        batch["vqa_scores"] = [[1.0] for _ in range(B)]
    batch = to_device(batch, dev, non_blocking=False)

    def step():
        for p in model.parameters():
            p.grad = None
        out = model(batch)
        loss = sum(v for k, v in out.items() if "loss" in k)
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    v = B / ms * 1e3
    return {"value": v, "unit": "pairs/s", "ms_per_step": ms, "per_gpu_batch": B, "image_size": R, "text_len": L,
            "tasks": tasks, "steps": steps, "warmup": warmup, "tflops": v * flops / 1e12, "last_loss": float(loss.detach()),
            "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}


def measure_fg_backbone(dev, B=8, Hi=800, Wi=1344, L=256, steps=3, warmup=2):
    """BASELINE configs[4] (fine-grained grounding, 800 px short side, 8 images / GPU), the part of it that is this repo's
    path: forward + backward of the FUSED BACKBONE (fiber_b200.modules.fusion_swin_fg — Swin-B over
    800 x 1344 images with padded 12 x 12 windows, roberta-base over 256 query tokens, 6 fused pairs) under a probe loss on
    its outputs; FPN / DyHead / detection losses are not part of it.  Next to it: the UNMODIFIED reference modules
    (baseline/ref_fg.py) in PyTorch eager under bf16 autocast on the same GPU, same inputs and loss."""
    import gc
    import types
    from fiber_b200.modules import fusion_swin_fg as M
    g = torch.Generator(device="cpu").manual_seed(77)
    img = torch.randn(B, 3, Hi, Wi, generator=g).to(dev)
    ids = torch.randint(3, 50265, (B, L), generator=g)
    ids[:, 0] = 0
    mask = torch.ones(B, L, dtype=torch.long)
    for b in range(B):  # grounding captions are short: 20 .. 60 real tokens of the 256-token query
        n = int(torch.randint(20, 60, (1,), generator=g))
        ids[b, n - 1] = 2
        ids[b, n:] = 1
        mask[b, n:] = 0
    tok = {"input_ids": ids.to(dev), "attention_mask": mask.to(dev)}

    def timed(model, autocast):
        def step():
            for p in model.parameters():
                p.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                outs, lang, _ = model(tok, types.SimpleNamespace(tensors=img))
            loss = sum(o.float().mean() for o in outs) + lang["hidden"].float().mean()
            loss.backward()
            return loss
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, float(loss.detach())

    torch.manual_seed(1234)
    torch.cuda.reset_peak_memory_stats()
    model = M.FusionSwinTransformer(M.SwinTransformer(drop_path_rate=0.2)).to(dev).train()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(("alpha_i2t", "alpha_t2i")):
                p.fill_(0.5)
    ms, loss = timed(model, False)
    out = {"value": B / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "per_gpu_batch": B, "image": [Hi, Wi], "text_len": L,
           "what": "fused backbone fwd+bwd (Swin-B + roberta-base, 6 fused pairs), probe loss; FPN / DyHead not included",
           "steps": steps, "warmup": warmup, "last_loss": loss,
           "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    del model
    gc.collect()
    torch.cuda.empty_cache()
    try:
        from baseline import ref_fg
        if not ref_fg.available():
            raise RuntimeError("baseline/_ref/fine_grained not installed")
        torch.manual_seed(1234)
        torch.cuda.reset_peak_memory_stats()
        ref, _swin, _rob = ref_fg.build(drop_path_rate=0.2)
        ref = ref.to(dev).train()
        with torch.no_grad():
            for n, p in ref.named_parameters():
                if n.endswith(("alpha_i2t", "alpha_t2i")):
                    p.fill_(0.5)
        rms, rloss = timed(ref, True)
        out["eager_reference"] = {"value": B / rms * 1e3, "unit": "images/s", "ms_per_step": rms, "dtype": "bf16 autocast",
                                  "impl": "unmodified reference modules (baseline/_ref/fine_grained), PyTorch eager",
                                  "last_loss": rloss, "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                                  "ours_over_eager": rms / ms}
        del ref
    except Exception as e:  # noqa: BLE001
        out["eager_reference"] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    gc.collect()
    torch.cuda.empty_cache()
    return out


def to_device(batch, dev, non_blocking=True):
    out = {}
    for k, v in batch.items():
        if isinstance(v, list) and v and torch.is_tensor(v[0]):
            out[k] = [t.to(dev, non_blocking=non_blocking) for t in v]
        elif torch.is_tensor(v):
            out[k] = v.to(dev, non_blocking=non_blocking)
        else:
            out[k] = v
    return out


def pin(batch):
    out = {}
    for k, v in batch.items():
        if isinstance(v, list) and v and torch.is_tensor(v[0]):
            out[k] = [t.pin_memory() for t in v]
        elif torch.is_tensor(v):
            out[k] = v.pin_memory()
        else:
            out[k] = v
    return out


def batch_bytes(batch):
    n = 0
    for v in batch.values():
        if isinstance(v, list) and v and torch.is_tensor(v[0]):
            n += sum(t.numel() * t.element_size() for t in v)
        elif torch.is_tensor(v):
            n += v.numel() * v.element_size()
    return n


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the UNMODIFIED reference (baseline/_ref, driven through its own
# FIBERTransformerSS.training_step by baseline/ref_harness.py) on the host cores; the oracle port of the
# same algorithm only when the reference package is not installed next to this file
# ---------------------------------------------------------------------------------------------
def fill_queues(model, seed=4321):
    """Steady-state ITC queues (what every run reaches after 4096 / global-batch steps, fiber_module.py:181-222):
    unit-norm feature columns, valid token ids, queue_total = queue_size — so that every N and both arms time the
    same work (hard negatives drawn from batch + a full queue)."""
    if not hasattr(model, "image_queue"):
        return
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for q in (model.image_queue, model.text_queue):
            v = torch.randn(q.shape, generator=g)
            q.copy_((v / v.norm(dim=0, keepdim=True)).to(q.device))
        n, L = model.text_input_queue.shape
        ids = torch.randint(3, 50264, (n, L), generator=g)
        ids[:, 0], ids[:, -1] = 0, 2
        model.text_input_queue.copy_(ids.to(model.text_input_queue.device))
        model.text_input_mask_queue.fill_(1)
        model.queue_total.fill_(n)
        model.queue_ptr.fill_(0)


def reference_available():
    try:
        from baseline import ref_harness
        return ref_harness.available()
    except Exception:  # noqa: BLE001
        return False


def ref_cpu_step_seconds(B, image_size, L, steps, warmup, threads):
    """fwd+bwd of the ITM+ITC+MLM step of the unmodified reference (fp32, CPU, all host threads)."""
    from baseline import ref_harness
    torch.set_num_threads(threads)
    st = ref_harness.RefStep(["itm", "itc", "mlm"], image_size, L, "cpu")
    batch = make_batch(B, image_size, L, seed=1234)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        st(batch)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    st.close()
    return sum(times) / len(times)


def cpu_baseline_record(B, image_size, L, steps, warmup):
    threads = os.cpu_count() or 1
    if reference_available():
        sec = ref_cpu_step_seconds(B, image_size, L, steps, warmup, threads)
        kind, what = "reference", "unmodified reference (baseline/_ref, FIBERTransformerSS.training_step + backward)"
    else:
        sec = cpu_step_seconds(B, image_size, L, steps, warmup, threads)
        kind, what = "port", "oracle port of the reference algorithm (reference package not installed)"
    return sec, {"value": B / sec, "unit": "pairs/s", "cores": threads, "kind": kind,
                 "sample": "B=%d pairs/step, %d timed step(s) of the same ITM+ITC+MLM %dpx workload, fp32, torch %d threads; %s"
                           % (B, steps, image_size, threads, what)}


def cpu_step_seconds(B, image_size, L, steps, warmup, threads):
    """fwd+bwd of the ITM+ITC+MLM step through oracle/fiber_oracle.py (fp32, CPU)."""
    from oracle import fiber_oracle as O
    from oracle import synth
    torch.set_num_threads(threads)
    cfg = config(["itm", "itc", "mlm"], image_size, L)
    from fiber_b200.modules import FIBERTransformerSS
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in FIBERTransformerSS(cfg).state_dict().items()
              if not k.startswith("rank_output")}
    sd = {k: v.requires_grad_(True) for k, v in synth.synth_state_dict(shapes).items()}
    batch = synth.synth_batch(B, image_size, L, seed=1234)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        itc = O.compute_itc(sd, cfg, batch, 0)
        img_idx = torch.multinomial(itc["weights_t2i"] + 1e-9, 1).view(-1) if B > 1 else torch.zeros(B, dtype=torch.long)
        txt_idx = torch.multinomial(itc["weights_i2t"] + 1e-9, 1).view(-1) if B > 1 else torch.zeros(B, dtype=torch.long)
        itm = O.compute_itm_hardneg(sd, cfg, batch, batch["image"][0][img_idx], batch["text_ids"][txt_idx],
                                    batch["text_masks"][txt_idx])
        mlm = O.compute_mlm(sd, cfg, batch)
        loss = itc["itc_loss"] + itm["itm_loss"] + mlm["mlm_loss"]
        for v in sd.values():
            v.grad = None
        loss.backward()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def bench_config(args, world):
    """The `config` object, identical for both arms (what is measured, not how)."""
    return {"workload": WORKLOAD, "per_gpu_batch": args.batch, "global_batch": args.batch * world,
            "image_size": args.image_size, "text_len": args.text_len, "parallelism": "dp%d" % world,
            "itc_queue": "full (4096 entries, steady state)",
            "l2": "per-step activations (>10 GB) exceed the 126 MB L2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch  # bounded sample of the workload: B pairs per step (one step is a few seconds on 16 cores)
    sec, rec = cpu_baseline_record(B, args.image_size, args.text_len, args.steps, args.warmup)
    v = B / sec
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "reference_arm": "the reference's own CPU implementation of the path on all host cores (kind '%s'); each step is a "
                         "bounded sample of B=%d pairs of the workload" % (rec["kind"], B),
        "cpu_baseline": rec,
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def profile_gemms(step, batch):
    """One extra (untimed) step with every tcgen05 GEMM launch bracketed by CUDA events on its stream:
    the dominant kernel's achieved TFLOP/s = sum(flops) / sum(duration), plus the per-shape breakdown."""
    from fiber_b200 import kernels as K
    K.GEMM_PROFILE = []
    try:
        step(batch)
        torch.cuda.synchronize()
        rec = K.GEMM_PROFILE
    finally:
        K.GEMM_PROFILE = None
    by = {}
    for key, fl, nb, e0, e1 in rec:
        d = by.setdefault(key, [0, 0.0, 0.0, 0.0])
        d[0] += 1; d[1] += fl; d[2] += nb; d[3] += e0.elapsed_time(e1)
    tot_ms = sum(d[3] for d in by.values())
    tot_fl = sum(d[1] for d in by.values())
    ranked = sorted(by.items(), key=lambda kv: -kv[1][3])
    dump = os.environ.get("FIBER_BENCH_DUMP")  # full per-shape table (tools / profiles), not part of the JSON line
    if dump:
        with open(dump, "w") as f:
            for k, d in ranked:
                f.write("%-44s n=%4d ms=%8.3f tflops=%7.1f gbs=%6.0f\n" % (str(k), d[0], d[3], d[1] / d[3] / 1e9, d[2] / d[3] / 1e6))
    top = ranked[:12]
    return {"launches": len(rec), "ms": tot_ms, "tflops": tot_fl / tot_ms / 1e9,
            "gbs": sum(d[2] for d in by.values()) / tot_ms / 1e6,
            "top": [{"mnk_epi": list(k), "n": d[0], "ms": round(d[3], 3), "tflops": round(d[1] / d[3] / 1e9, 1),
                     "gbs": round(d[2] / d[3] / 1e6)} for k, d in top]}


def gpu_speed_probe(dev, world):
    """How uniform are the GPUs of this job?  Every rank times the same 20 launches of the step's largest GEMM
    (147456 x 512 x 2048, no communication) after the timed region; the per-rank TFLOP/s are gathered.  A data-parallel
    step runs at the pace of its slowest rank (the ranks meet in the ITC all_gathers and in the gradient all-reduce),
    so the spread printed here is a floor for the scaling loss that no overlap scheme can recover."""
    from fiber_b200 import kernels as K
    a = torch.randn(147456, 2048, device=dev).to(torch.bfloat16)
    b = torch.randn(512, 2048, device=dev).to(torch.bfloat16)
    for _ in range(5):
        K.gemm(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 40
    for _ in range(n):
        K.gemm(a, b)
    e1.record()
    torch.cuda.synchronize()
    tf = 2.0 * 147456 * 512 * 2048 * n / (e0.elapsed_time(e1) / 1e3) / 1e12
    if world == 1:
        return {"gemm_tflops_per_rank": [round(tf, 1)]}
    t = torch.tensor([tf], device=dev)
    out = torch.empty(world, device=dev)
    torch.distributed.all_gather_into_tensor(out, t)
    v = [round(float(x), 1) for x in out.tolist()]
    return {"gemm_tflops_per_rank": v, "slowest_over_fastest": round(min(v) / max(v), 4),
            "slowest_over_rank0": round(min(v) / v[0], 4)}


def profile_timeline(step, batch, out_path):
    """One more (untimed) step under torch.profiler on every rank (the step contains collectives); the rank that is
    given a path writes the kernel timeline summary: totals per kernel, NCCL kernel spans, busy / idle time of the
    compute stream and of the union of all streams."""
    from collections import defaultdict
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        step(batch)
        torch.cuda.synchronize()
    if out_path is None:
        return
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    t0, t1 = ks[0][0], max(k[1] for k in ks)

    def union(iv):
        busy, end = 0.0, None
        for s_, e_ in sorted(iv):
            if end is None or s_ > end:
                busy += e_ - s_
                end = e_
            elif e_ > end:
                busy += e_ - end
                end = e_
        return busy

    nccl = [(s_, e_, n) for s_, e_, n in ks if "nccl" in n.lower()]
    comp = [(s_, e_, n) for s_, e_, n in ks if "nccl" not in n.lower()]
    agg = defaultdict(lambda: [0, 0.0])
    for s_, e_, n in ks:
        n = n.split("(")[0].replace("void ", "")
        agg[n][0] += 1
        agg[n][1] += e_ - s_
    lines = ["span %.3f ms, %d kernels; compute kernels busy %.3f ms (sum %.3f ms), NCCL kernels busy %.3f ms in %d launches, "
             "all streams busy %.3f ms" % ((t1 - t0) / 1e3, len(ks), union([(a, b) for a, b, _ in comp]) / 1e3,
                                           sum(b - a for a, b, _ in comp) / 1e3, union([(a, b) for a, b, _ in nccl]) / 1e3,
                                           len(nccl), union([(a, b) for a, b, _ in ks]) / 1e3)]
    lines.append("NCCL kernels (start ms, duration ms, name):")
    for s_, e_, n in nccl:
        if e_ - s_ > 50:
            lines.append("  %8.2f %8.3f  %s" % ((s_ - t0) / 1e3, (e_ - s_) / 1e3, n[:90]))
    lines.append("%-86s %6s %10s %9s" % ("kernel", "count", "ms", "avg us"))
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        lines.append("%-86s %6d %10.3f %9.1f" % (n[:86], c, us / 1e3, us / c))
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")


def run_ours(args):
    from fiber_b200 import lib, ops
    from fiber_b200.modules import FIBERTransformerSS, fiber_utils

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: ONE JSON line is the contract
        torch.distributed.init_process_group("nccl", device_id=dev)
    lib.check(lib.load().fiber_init(), "init")
    B, R, L = args.batch, args.image_size, args.text_len
    cfg = config(["itm", "itc", "mlm"], R, L)
    torch.manual_seed(1234)
    model = FIBERTransformerSS(cfg).to(dev)
    with torch.no_grad():  # alpha gates at 0.5 so the cross-attention branches carry signal
        for n, p in model.named_parameters():
            if n.endswith(("alpha_i2t", "alpha_t2i")):
                p.fill_(0.5)
    fill_queues(model)  # steady-state ITC queues: every N and both arms time the same work
    model.train()
    fiber_utils.set_task(model)
    ops.set_dropout_seed(1234 + rank)
    step_model = model
    if world > 1:
        # the set of grad-less parameters is fixed per task mix => static graph
        ddp_kw = dict(static_graph=True) if os.environ.get("FIBER_DDP_STATIC", "1") == "1" else \
            dict(find_unused_parameters=True)
        # broadcast_buffers=False: DDP's default re-broadcasts every module buffer from rank 0 at each forward, and this
        # model's buffers include the 4096-deep ITC queues (image_input_queue alone is 4096 x 3 x 384 x 384 fp32 = 7.2 GB
        # -> ~18-20 ms of flatten + NCCL broadcast + unflatten per step).  The queues are identical on every rank by
        # construction (DDP syncs module state once at construction, and every rank applies the same all-gathered
        # update, fiber_module.py:181-222; tests/test_dist_gloo.py), there is no BatchNorm, so the broadcast moves
        # bytes without changing a value.  FIBER_DDP_BCAST_BUFFERS=1 restores the default.
        step_model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                               bucket_cap_mb=int(os.environ.get("FIBER_DDP_BUCKET_MB", "100")),
                                                               broadcast_buffers=os.environ.get("FIBER_DDP_BCAST_BUFFERS", "0") == "1",
                                                               **ddp_kw)
        if os.environ.get("FIBER_DDP_BF16", "0") == "1":
            # opt-in: gradient buckets travel as bf16 (0.56 GB instead of 1.13 GB per step).  Measured at N = 8
            # (profiles/r2_scaling_ab.txt): 2386 pairs/s with, 2388 without — the fp32 all-reduce is already hidden
            # behind the backward, so the default keeps the reference's fp32 gradient averaging
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            step_model.register_comm_hook(None, default_hooks.bf16_compress_hook)
    host = pin(make_batch(B, R, L, seed=1234 + rank))
    h2d = batch_bytes(host)

    # the heads on top of infer() (callers, plain torch fp32 modules) may use TF32 tensor cores
    torch.backends.cuda.matmul.allow_tf32 = True

    def step(batch):
        # clear the gradients first: at this point the GPU is still working through the previous step's backward, so
        # the 754-parameter host loop is hidden; between forward and backward it showed up as GPU idle time
        # (profiles/r1_step_timeline.txt: ~1.9 ms gap at the forward / backward boundary)
        for p in model.parameters():
            p.grad = None
        out = step_model(batch)
        loss = sum(v for k, v in out.items() if "loss" in k)
        loss.backward()
        return loss

    def timed(n, from_host):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.launch_count()
        e0.record()
        res = None
        for _ in range(n):
            b = to_device(host, dev) if from_host else dev_batch
            loss = step(b)
            if from_host:
                res = loss.item()  # device -> host read of the step's result
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms, lib.launch_count() - l0, res

    dev_batch = to_device(host, dev)
    timed(max(args.warmup, 3), False)
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches, _ = timed(args.steps, False)
    clocks = sampler.stop()
    ms_e2e, _, last_loss = timed(args.steps, True)
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
    gemm_stats = profile_gemms(step, dev_batch)  # every rank runs it: the step contains DDP collectives
    rank_probe = gpu_speed_probe(dev, world)
    # the same step followed by the optimizer (fused multi-tensor AdamW, one launch; the bf16 weight
    # copies are re-cast by the next forward) — reported next to the fwd+bwd metric, not instead of it
    (optimizer,), (sched,) = fiber_utils.set_schedule(model)

    def full_step(batch):
        loss = step(batch)
        optimizer.step()
        sched["scheduler"].step()
        return loss

    for _ in range(3):  # the first steps allocate the AdamW state and re-shape the allocator's pools (one 250-ms step)
        full_step(dev_batch)
    torch.cuda.synchronize()
    n_opt = max(3, args.steps // 2)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_opt + 1)]
    evs[0].record()
    for i in range(n_opt):
        full_step(dev_batch)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ms_opt_each = [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(n_opt)]
    ms_opt = sorted(ms_opt_each)[n_opt // 2]  # median: every step is listed in ms_each
    del optimizer
    if args.profile_out:
        profile_timeline(step, dev_batch, args.profile_out if rank == 0 else None)

    if rank != 0:
        return
    value = args.steps * B * world / (ms / 1e3)
    e2e = args.steps * B * world / (ms_e2e / 1e3)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = value / world * FLOPS_PER_PAIR / 1e12
    hbm = peaks.get("hbm_gbs", 6650.0)
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": bench_config(args, world),
        "value_per_gpu": value / world,
        "run_info": {"last_loss": last_loss, "peak_mem_gib": round(peak_mem, 1), "kernel_options": _kernel_options(),
                     "gpu_speed_probe": rank_probe,
                     "with_optimizer": {"ms_per_step": ms_opt, "ms_each": ms_opt_each, "pairs_per_s": B * world / ms_opt * 1e3, "steps": n_opt,
                                        "optimizer": "fiber_b200.optim.FusedAdamW (HF AdamW order, one launch) + LR schedule"},
                     "note": "`value` is the whole-job aggregate over n_gpus (contract); value_per_gpu is the metric's "
                             "per-GPU figure"},
        "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clocks,
        # dominant kernel = the tcgen05 GEMM (all of its launches in one step, CUDA events on the launching
        # stream, one extra untimed step): achieved = sum(2MNK) / sum(duration)
        "roofline": {"bound": "tensor", "achieved": gemm_stats["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": gemm_stats["tflops"] / peak_tf, "traffic": GEMM_NCU_TRAFFIC["dram_bytes"],
                     "traffic_note": GEMM_NCU_TRAFFIC["note"],
                     "kernel": "fiber::gemm_tcgen05_kernel", "launches_per_step": gemm_stats["launches"],
                     "kernel_ms_per_step": round(gemm_stats["ms"], 2),
                     "algorithmic_gbs": round(gemm_stats["gbs"]), "hbm_frac": gemm_stats["gbs"] / hbm,
                     "note": "achieved = sum(2MNK) / sum(duration) over ALL GEMM launches of one step, CUDA events on "
                             "the launching stream; peak = %s bf16 sustained (kernel timed inside a long step); "
                             "the kernel is HBM-bound on the K<=256 shapes of Swin stages 0/1 (hbm_frac = algorithmic "
                             "bytes / HBM peak)" % ("measured" if peaks else "fallback")},
        "step_roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                          "frac": achieved / peak_tf,
                          "note": "whole step per GPU: pairs/s x 1636.23 GFLOP/pair"},
        "gemm_breakdown": gemm_stats["top"],
    }
    if world == 1 and (args.eager_baseline or args.extra_configs):
        del step_model, model, dev_batch
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    if world == 1 and args.extra_configs:
        # BASELINE.json configs[2] / [3] on the same kernels (extra keys, not the headline)
        line["extra_configs"] = {}
        for name in EXTRA_CONFIGS:
            try:
                torch.cuda.reset_peak_memory_stats()
                line["extra_configs"][name] = measure_extra_config(name, dev)
            except Exception as e:  # noqa: BLE001  (never let an extra break the headline line)
                line["extra_configs"][name] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
            import gc
            gc.collect()
            torch.cuda.empty_cache()
        try:  # configs[4]: the fused backbone of the fine-grained model
            line["extra_configs"]["fg800_backbone"] = measure_fg_backbone(dev)
        except Exception as e:  # noqa: BLE001
            line["extra_configs"]["fg800_backbone"] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if world == 1 and args.extra_configs:
        try:  # the image side of the input pipeline (ragged uint8 images -> batch["image"][0])
            from tools.bench_image import measure_image_pipeline
            line["extra_configs"]["input_pipeline_albef384"] = measure_image_pipeline(dev, steps=10, warmup=3)
        except Exception as e:  # noqa: BLE001
            line["extra_configs"]["input_pipeline_albef384"] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if world == 1 and args.eager_baseline:
        # the north-star's denominator: the reference PyTorch-eager GPU path, timed in this process after our arm
        line["eager_gpu_baseline"] = eager_gpu_baseline(args, dev, host, value)
    if args.cpu_baseline and world == 1:
        _, line["cpu_baseline"] = cpu_baseline_record(args.cpu_batch, R, L, 1, 0)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def eager_gpu_baseline(args, dev, host, our_value):
    """The UNMODIFIED reference (baseline/_ref) as PyTorch eager on this GPU under bf16 autocast — same workload, same
    synthetic batch schema, full ITC queue, dropout / DropPath on, optimizer excluded — at the largest per-GPU batch of
    64 / 32 / 16 / 8 that fits (the reference materialises every score matrix; SURVEY.md §7 'Eager baseline at B=64')."""
    import gc
    if not reference_available():
        return {"unavailable": "reference package not installed (baseline/_ref missing)"}
    from baseline import ref_harness
    R, L = args.image_size, args.text_len
    tried = []
    for B in (args.batch, 32, 16, 8):
        if B > args.batch or B in tried:
            continue
        tried.append(B)
        gc.collect()
        torch.cuda.empty_cache()
        st = None
        try:
            st = ref_harness.RefStep(["itm", "itc", "mlm"], R, L, dev, autocast=torch.bfloat16)
            batch = to_device(make_batch(B, R, L, seed=1234), dev, non_blocking=False)
            for _ in range(2):
                st(batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 3
            e0.record()
            for _ in range(n):
                loss = st(batch)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            v = B / ms * 1e3
            mem = torch.cuda.max_memory_allocated() / 2 ** 30
            st.close()
            return {"value": v, "unit": "pairs/s", "per_gpu_batch": B, "ms_per_step": ms, "dtype": "bf16 autocast",
                    "impl": "unmodified reference (baseline/_ref), FIBERTransformerSS.training_step + backward, PyTorch eager",
                    "steps": n, "warmup": 2, "last_loss": float(loss), "peak_mem_gib": round(mem, 1),
                    "batches_tried": tried, "ours_over_eager": our_value / v}
        except torch.cuda.OutOfMemoryError:
            if st is not None:
                st.close()
            st = None
            continue
        except Exception as e:  # noqa: BLE001  (a label must never break the measurement)
            if st is not None:
                st.close()
            return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:300]), "batches_tried": tried}
    return {"unavailable": "out of memory at every batch size tried", "batches_tried": tried}


def _kernel_options():
    """Kernel selections in effect (fiber_set_option / FIBER_* environment)."""
    try:
        from fiber_b200 import lib
        from fiber_b200 import ops
        opts = {k: lib.get_option(k) for k in ("winattn_tc", "attn_small", "attn_sk", "gemm_cta2", "pdl")}
        from fiber_b200.modules import objectives
        opts["mlm_fused_ce"] = int(objectives.FUSED_MLM_CE)
        opts["gelu_cache"] = int(ops.GELU_CACHE)
        opts["gelu_onepass"] = int(ops.GELU_ONEPASS)
        opts["gelu_grad_prefetch"] = int(ops.GELU_GRAD_PREFETCH)
        from fiber_b200 import kernels
        opts["res_prefetch"] = int(kernels.RES_PREFETCH)
        opts["ddp_bf16_buckets"] = int(os.environ.get("FIBER_DDP_BF16", "0") == "1")
        opts["itc_async_queue"] = int(os.environ.get("FIBER_ITC_ASYNC_QUEUE", "0") == "1")
        return opts
    except Exception as e:  # never let a label break the measurement
        return {"error": str(e)}


def main():
    # NCCL prints its version banner (NCCL_DEBUG >= VERSION) to stdout by default; ONE JSON line on stdout is the contract
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="pairs per GPU")
    ap.add_argument("--image-size", type=int, default=384)
    ap.add_argument("--text-len", type=int, default=40)
    ap.add_argument("--cpu-batch", type=int, default=2, help="pairs per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-eager-baseline", dest="eager_baseline", action="store_false")
    ap.add_argument("--no-extra-configs", dest="extra_configs", action="store_false",
                    help="skip the VQA-576 / ITC-only extra measurements of the N = 1 line")
    ap.add_argument("--profile-out", default="", help="write a torch.profiler kernel-timeline summary of one extra step here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
