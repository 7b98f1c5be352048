"""ORACLE support — deterministic synthetic weights and batches (test infrastructure).

Weights are never shared by RNG seed of the reference initialisers (timm / HF init order is not
reproduced); instead every tensor of a state_dict is filled from its own CPU generator seeded by
crc32(name), so the reference (in tools/make_golden.py), the oracle and the CUDA path all get
bit-identical parameters wherever they run (same torch build => same CPU Philox/mt19937 stream).
Batches follow SURVEY.md §8(d).
"""
import zlib

import torch


def _gen(name, salt=0):
    return torch.Generator(device="cpu").manual_seed((zlib.crc32(name.encode()) + 7919 * salt) & 0x7FFFFFFF)


def synth_tensor(name, shape, dtype=torch.float32, salt=0):
    """Value recipe by parameter role (chosen so every branch of the path carries signal)."""
    g = _gen(name, salt)
    leaf = name.split(".")[-1]
    if name.endswith(("alpha_i2t", "alpha_t2i")):
        return torch.full(shape, 0.5, dtype=dtype)  # gates init to 0 in the reference: a wrong
        # cross-attention would pass unnoticed, so parity runs use 0.5
    if name == "temp":
        return torch.full(shape, 0.07, dtype=dtype)
    if leaf == "relative_position_bias_table":
        return (torch.randn(shape, generator=g) * 0.2).to(dtype)
    is_norm = any(k in name for k in ("norm", "LayerNorm", "vqa_classifier.1"))
    if is_norm and leaf == "weight":
        return (1.0 + 0.1 * torch.randn(shape, generator=g)).to(dtype)
    if is_norm and leaf == "bias":
        return (0.1 * torch.randn(shape, generator=g)).to(dtype)
    if leaf == "bias":
        return (0.02 * torch.randn(shape, generator=g)).to(dtype)
    if name.endswith("_queue") and "input" not in name:
        q = torch.randn(shape, generator=g)
        return (q / q.norm(dim=0, keepdim=True)).to(dtype)
    if len(shape) >= 2:
        # reference initialiser scale (trunc_normal / normal std 0.02); inputs ("in.*") get unit scale
        std = 1.0 if name.startswith(("in.", "probe.")) else 0.02
        return (torch.randn(shape, generator=g) * std).to(dtype)
    return (0.02 * torch.randn(shape, generator=g)).to(dtype)


def synth_state_dict(shapes, salt=0, skip=("relative_position_index", "attn_mask", "position_ids", "queue_ptr",
                                            "queue_total", "image_input_queue", "text_input_queue",
                                            "text_input_mask_queue")):
    """shapes: {name: (shape, dtype)} -> {name: tensor}.  Structural buffers are skipped."""
    out = {}
    for name in sorted(shapes):
        shape, dtype = shapes[name]
        if name.split(".")[-1] in skip or not dtype.is_floating_point:
            continue
        out[name] = synth_tensor(name, tuple(shape), dtype, salt)
    return out


def synth_batch(B, image_size, L, seed=1234, vocab=50265, false_image=False, vqa=False):
    """Synthetic batch dict with the schema of base_dataset.collate (datasets/base_dataset.py:172-245)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    batch = {"image": [torch.randn(B, 3, image_size, image_size, generator=g)]}
    if false_image:
        batch["false_image_0"] = [torch.randn(B, 3, image_size, image_size, generator=g)]
    ids = torch.randint(3, vocab - 1, (B, L), generator=g)
    masks = torch.ones(B, L, dtype=torch.long)
    lens = torch.randint(8, L + 1, (B,), generator=g)
    for b in range(B):
        n = int(lens[b]) if b % 2 == 1 else L  # every other row is padded
        ids[b, 0] = 0
        ids[b, n - 1] = 2
        ids[b, n:] = 1
        masks[b, n:] = 0
    batch["text_ids"] = ids
    batch["text_masks"] = masks
    batch["text_labels"] = torch.full((B, L), -100)
    pick = (torch.rand(B, L, generator=g) < 0.15) & (ids > 2)
    pick[:, 1] = True  # at least one masked position per row
    batch["text_ids_mlm"] = torch.where(pick, torch.full_like(ids, vocab - 1), ids)
    batch["text_labels_mlm"] = torch.where(pick, ids, torch.full_like(ids, -100))
    batch["text"] = ["synthetic caption"] * B
    if vqa:
        batch["vqa_labels"] = [[int(torch.randint(0, 3129, (1,), generator=g))] for _ in range(B)]
        batch["vqa_scores"] = [[1.0] for _ in range(B)]
    return batch
