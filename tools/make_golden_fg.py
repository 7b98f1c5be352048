"""Golden vectors for the FINE-GRAINED fused backbone from the UNMODIFIED reference modules:
fine_grained/maskrcnn_benchmark/modeling/backbone/fusion_swin_transformer_v2.py (FusionSwinTransformer.forward run as
is) over language_backbone/roberta_fused_model_v2.py, imported by path in the build container under the shims of
baseline/ref_shims.py (+ two HF 4.6 -> 5.x aliases).  Weights / inputs are oracle.synth recipes (crc32-seeded by name),
so the oracle and the CUDA path rebuild the same tensors.  Writes tests/golden/fg_fused_backbone.pt.
    python tools/make_golden_fg.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_shims  # noqa: E402
from oracle import synth  # noqa: E402

NS = types.SimpleNamespace


def build():
    from baseline import ref_fg
    fusion, swin, rob = ref_fg.build(drop_path_rate=0.0)
    # synthetic weights by (coarse-oracle style) name
    shapes = {}
    for k, v in swin.state_dict().items():
        shapes["vit_model." + k] = (tuple(v.shape), v.dtype)
    for k, v in rob.state_dict().items():
        shapes["text_transformer." + k] = (tuple(v.shape), v.dtype)
    sd = synth.synth_state_dict(shapes)
    swin.load_state_dict({k[len("vit_model."):]: v for k, v in sd.items() if k.startswith("vit_model.")}, strict=False)
    rob.load_state_dict({k[len("text_transformer."):]: v for k, v in sd.items() if k.startswith("text_transformer.")},
                        strict=False)
    return fusion.eval(), sd, shapes


def inputs(B, Hi, Wi, L, seed=4242):
    g = torch.Generator().manual_seed(seed)
    img = synth.synth_tensor("in.fg_img", (B, 3, Hi, Wi))
    ids = torch.randint(3, 50265, (B, L), generator=g)
    ids[:, 0] = 0
    mask = torch.ones(B, L, dtype=torch.long)
    for b in range(1, B):  # ragged lengths: pad id 1 behind an </s>
        n = int(torch.randint(L // 2, L, (1,), generator=g))
        ids[b, n - 1] = 2
        ids[b, n:] = 1
        mask[b, n:] = 0
    return img, ids, mask


def main():
    fusion, sd, shapes = build()
    gold = {"cases": {}}
    for name, (B, Hi, Wi, L) in {"pad_224x320": (2, 224, 320, 20), "nopad_384x384": (1, 384, 384, 12)}.items():
        img, ids, mask = inputs(B, Hi, Wi, L)
        for p in fusion.parameters():
            p.grad = None
        vis, lang, _ = fusion({"input_ids": ids, "attention_mask": mask}, NS(tensors=img))
        # scalar probe loss: fixed random projections of every output (oracle.synth recipe "probe.*")
        loss = sum((v * synth.synth_tensor("probe.fg.%s.%d" % (name, i), tuple(v.shape))).mean() for i, v in enumerate(vis))
        loss = loss + (lang["hidden"] * synth.synth_tensor("probe.fg.%s.t" % name, tuple(lang["hidden"].shape))).mean()
        loss.backward()
        grads = {}
        for mod, pre in ((fusion.backbone.body, "vit_model."), (fusion.language_backbone.body.model, "text_transformer.")):
            for n, p in mod.named_parameters():
                if p.grad is not None:
                    grads[pre + n] = float(p.grad.double().norm())
        gold["cases"][name] = {
            "B": B, "Hi": Hi, "Wi": Wi, "L": L, "ids": ids, "mask": mask,
            "vis_shapes": [tuple(v.shape) for v in vis],
            "vis_sample": [v.detach()[:, ::7, ::3, ::3].clone() for v in vis],
            "vis_norm": [float(v.detach().double().norm()) for v in vis],
            "hidden": lang["hidden"].detach().clone(), "aggregate": lang["aggregate"].detach().clone(),
            "loss": float(loss), "grad_norms": grads,
        }
        print(name, [tuple(v.shape) for v in vis], "loss", float(loss), len(grads), "grads")
    gold["state_keys"] = {k: (tuple(s), str(d)) for k, (s, d) in shapes.items()}
    out = os.path.join(ROOT, "tests", "golden", "fg_fused_backbone.pt")
    torch.save(gold, out)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
