"""Context number: the reference ALGORITHM run as PyTorch eager on the GPU (oracle restatement, bf16
autocast and fp32) — the 'reference PyTorch-eager GPU path' the north-star's 4x target refers to.
The unmodified reference cannot travel to the GPU box (needs pytorch_lightning/timm/sacred), so
this uses oracle/fiber_oracle.py, which issues the same eager ops (Linear, bmm, softmax, LayerNorm,
roll-free gathers).  Not part of bench.py's contract; results are recorded in profiles/."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import fiber_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402
from fiber_b200.modules import FIBERTransformerSS  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    dev = torch.device("cuda:0")
    cfg = bench.config(["itm", "itc", "mlm"], 384, 40)
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in FIBERTransformerSS(cfg).state_dict().items()
              if not k.startswith("rank_output")}
    sd = {k: v.to(dev).requires_grad_(True) for k, v in synth.synth_state_dict(shapes).items()}
    batch = bench.to_device(synth.synth_batch(B, 384, 40, seed=1234), dev, non_blocking=False)
    torch.backends.cuda.matmul.allow_tf32 = True

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == "bf16")):
            itc = O.compute_itc(sd, cfg, batch, 0)
            img_idx = torch.multinomial(itc["weights_t2i"].float() + 1e-9, 1).view(-1)
            txt_idx = torch.multinomial(itc["weights_i2t"].float() + 1e-9, 1).view(-1)
            itm = O.compute_itm_hardneg(sd, cfg, batch, batch["image"][0][img_idx], batch["text_ids"][txt_idx],
                                        batch["text_masks"][txt_idx])
            mlm = O.compute_mlm(sd, cfg, batch)
            loss = itc["itc_loss"] + itm["itm_loss"] + mlm["mlm_loss"]
        for v in sd.values():
            v.grad = None
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"impl": "oracle-eager-gpu", "dtype": dtype, "per_gpu_batch": B, "ms_per_step": ms,
                      "pairs_per_s": B / ms * 1e3, "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
