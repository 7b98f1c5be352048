"""Print parity numbers of the CUDA path vs the fp32 oracle (run on the GPU box)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fiber_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402
from fiber_b200.modules import FIBERTransformerSS  # noqa: E402
from fiber_b200 import lib  # noqa: E402


def cfg_for(tasks, image_size, L=40):
    loss_names = {"itm": 0, "mlm": 0, "itc": 0, "vqa": 0, "nlvr2": 0, "caption_mle": 0, "caption_gold": 0, "caption_cider": 0}
    loss_names.update({t: 1 for t in tasks})
    return dict(loss_names=loss_names, image_size=image_size, vit="swin_base_patch4_window12_384_in22k",
                input_image_embed_size=1024, input_text_embed_size=768, pretrained_vit=False, vqav2_label_size=3129,
                max_text_len=L, tokenizer="roberta-base", vocab_size=50265, hidden_size=768, num_heads=12,
                num_layers=12, mlp_ratio=4, drop_rate=0.1, num_fuse_block=6, itc_pooler=True, load_path="",
                test_only=False, optim_type="adamw", learning_rate=1e-5, weight_decay=0.01, decay_power=1,
                max_steps=100000, warmup_steps=10000, end_lr=0, lr_mult_head=5, lr_mult_cross_modal=5)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item(), ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def main():
    dev = torch.device("cuda:0")
    lib.check(lib.load().fiber_init(), "init")
    image_size = int(sys.argv[1]) if len(sys.argv) > 1 else 224
    B, L = 2, 40
    cfg = cfg_for(["itm", "mlm"], image_size)
    model = FIBERTransformerSS(cfg)
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items() if not k.startswith("rank_output")}
    sd = synth.synth_state_dict(shapes)
    model.load_state_dict(sd, strict=False)
    model.to(dev)
    sd = {k: v.to(dev) for k, v in sd.items()}
    batch = synth.synth_batch(B, image_size, L, seed=1234, false_image=True)
    batch = {k: ([t.to(dev) for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else (v.to(dev) if torch.is_tensor(v) else v)) for k, v in batch.items()}
    model.eval()
    for mode in ("fused", "image_only", "text_only"):
        with torch.no_grad():
            t = time.time()
            r = model.infer(batch, image_only=(mode == "image_only"), text_only=(mode == "text_only"))
            torch.cuda.synchronize()
            t1 = time.time() - t
            o = O.infer(sd, cfg, batch["image"][0], batch["text_ids"], batch["text_masks"],
                        image_only=(mode == "image_only"), text_only=(mode == "text_only"))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ob = O.infer(sd, cfg, batch["image"][0], batch["text_ids"], batch["text_masks"],
                             image_only=(mode == "image_only"), text_only=(mode == "text_only"))
        for k in ("cls_feats", "text_feats", "image_feats"):
            if o[k] is not None:
                print("%-10s %-11s ours max/l2 rel %.4g %.4g | oracle-bf16-autocast %.4g %.4g   (%.1f ms)" %
                      ((mode, k) + rel(r[k], o[k]) + rel(ob[k], o[k]) + (t1 * 1e3,)))
    # training step (dropout off): loss + grads
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0
    g = torch.Generator().manual_seed(5)
    labels = torch.tensor([1.0, 0.0])[torch.randperm(2, generator=g)].to(dev)
    from fiber_b200.modules import objectives as OBJ
    model.zero_grad()
    model.current_tasks = ["mlm", "itm"]
    loss = OBJ.compute_mlm(model, batch)["mlm_loss"] + OBJ.compute_itm(model, batch, labels)["itm_loss"]
    loss.backward()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo = O.compute_mlm(sdg, cfg, batch)["mlm_loss"] + O.compute_itm(sdg, cfg, batch, labels)["itm_loss"]
    lo.backward()
    print("loss ours %.6f oracle %.6f" % (loss.item(), lo.item()))
    worst = []
    for n, p in model.named_parameters():
        if n.startswith("rank_output"):
            continue
        go = sdg[n].grad
        if p.grad is None or go is None:
            if (p.grad is None) != (go is None or float(go.abs().sum()) == 0):
                print("GRAD PRESENCE MISMATCH", n, p.grad is None, go is None)
            continue
        worst.append((rel(p.grad, go)[1], n, float(go.norm())))
    worst.sort(reverse=True)
    for w in worst[:25]:
        print("grad l2-rel %.4g  %-70s |g|=%.3g" % w)
    print("median grad l2-rel %.4g over %d params" % (worst[len(worst) // 2][0], len(worst)))
    print("launches", lib.launch_count())


if __name__ == "__main__":
    main()
