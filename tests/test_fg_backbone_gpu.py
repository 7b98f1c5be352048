"""Fine-grained fused backbone on the CUDA path: fiber_b200.modules.fusion_swin_fg against the oracle
(oracle/fiber_oracle_fg.py) on this GPU and against the unmodified reference's fixture (tests/golden/fg_fused_backbone.pt,
tools/make_golden_fg.py)."""
import os
import types

import pytest
import torch

from oracle import fiber_oracle_fg as FG
from oracle import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _l2rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def fg_model(cuda_dev):
    from fiber_b200.modules import fusion_swin_fg as M
    gold = torch.load(os.path.join(GOLD, "fg_fused_backbone.pt"), weights_only=False)
    dt = {"torch.float32": torch.float32, "torch.int64": torch.int64}
    sd = synth.synth_state_dict({k: (s, dt[d]) for k, (s, d) in gold["state_keys"].items()})
    model = M.FusionSwinTransformer(M.SwinTransformer(drop_path_rate=0.0))
    # the reference's own state_dict names: every parameter of the fixture must find its place
    model.load_reference_state({k[len("vit_model."):]: v for k, v in sd.items() if k.startswith("vit_model.")},
                               {k[len("text_transformer."):]: v for k, v in sd.items() if k.startswith("text_transformer.")})
    ours = {"vit_model." + k for k, _ in model.backbone.body.named_parameters()} | \
           {"text_transformer." + k for k, _ in model.language_backbone.body.model.named_parameters()}
    assert ours == set(sd), (sorted(ours - set(sd))[:5], sorted(set(sd) - ours)[:5])
    return model.to(cuda_dev).eval(), gold, {k: v.to(cuda_dev) for k, v in sd.items()}


@pytest.mark.parametrize("case", ["pad_224x320", "nopad_384x384"])
def test_fg_fused_backbone_vs_oracle_and_reference_fixture(cuda_dev, fg_model, case):
    model, gold, sd = fg_model
    c = gold["cases"][case]
    img = synth.synth_tensor("in.fg_img", (c["B"], 3, c["Hi"], c["Wi"])).to(cuda_dev)
    ids, mask = c["ids"].to(cuda_dev), c["mask"].to(cuda_dev)
    model.zero_grad()
    outs, lang, _ = model({"input_ids": ids, "attention_mask": mask}, types.SimpleNamespace(tensors=img))
    assert [tuple(o.shape) for o in outs] == c["vis_shapes"]
    # ... against the unmodified reference (fixture)
    for o, smp, nrm in zip(outs, c["vis_sample"], c["vis_norm"]):
        assert _l2rel(o.detach()[:, ::7, ::3, ::3].cpu(), smp) < 3e-2
        assert abs(float(o.detach().double().norm()) - nrm) < 2e-2 * nrm
    assert _l2rel(lang["hidden"].detach().cpu(), c["hidden"]) < 3e-2
    assert _l2rel(lang["aggregate"].detach().cpu(), c["aggregate"]) < 3e-2
    # ... against the fp32 oracle on this GPU, element-wise over the full maps, next to the oracle's own bf16-autocast error
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    r_outs, r_lang = FG.fused_backbone(sdg, img, ids, mask)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a_outs, a_lang = FG.fused_backbone(sd, img, ids, mask)
    for i, (o, r, a) in enumerate(zip(outs, r_outs, a_outs)):
        e_o, e_a = _l2rel(o.detach(), r.detach()), _l2rel(a, r.detach())
        print("%s stage %d: l2-rel ours %.3g / autocast %.3g" % (case, i + 2, e_o, e_a))
        assert e_o <= 2.0 * e_a + 1e-3, (i, e_o, e_a)
    e_o, e_a = _l2rel(lang["hidden"].detach(), r_lang["hidden"].detach()), _l2rel(a_lang["hidden"], r_lang["hidden"].detach())
    print("%s text: l2-rel ours %.3g / autocast %.3g" % (case, e_o, e_a))
    assert e_o <= 2.0 * e_a + 1e-3
    # ... gradients of the fixture's probe loss
    def probe(vs, hidden):
        l = sum((v * synth.synth_tensor("probe.fg.%s.%d" % (case, i), tuple(v.shape)).to(cuda_dev)).mean() for i, v in enumerate(vs))
        return l + (hidden * synth.synth_tensor("probe.fg.%s.t" % case, tuple(hidden.shape)).to(cuda_dev)).mean()
    loss, r_loss = probe(outs, lang["hidden"]), probe(r_outs, r_lang["hidden"])
    # the probe loss is a mean of ~1e6 signed products (|loss| ~ 1e-3 by cancellation): compare on the absolute scale
    assert abs(float(r_loss.detach()) - c["loss"]) < 2e-5 and abs(float(loss.detach()) - c["loss"]) < 3e-4
    loss.backward()
    r_loss.backward()
    errs = []
    named = [("vit_model." + n, p) for n, p in model.backbone.body.named_parameters()] + \
            [("text_transformer." + n, p) for n, p in model.language_backbone.body.model.named_parameters()]
    scale = max(float(v.grad.norm()) for v in sdg.values() if v.grad is not None)
    gate = max(float(sdg[n].grad.abs().max()) for n, p in named if p.numel() == 1 and sdg[n].grad is not None)
    for n, p in named:
        go = sdg[n].grad
        if go is None or float(go.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, "unexpected gradient for " + n
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        if float(go.norm()) < 1e-6 * scale or n.endswith("key.bias"):
            continue
        assert abs(float(go.double().norm()) - c["grad_norms"][n]) <= 5e-3 * c["grad_norms"][n] + 1e-9, n  # oracle == reference
        errs.append((float((p.grad - go).abs().max()) / gate if p.numel() == 1 else _l2rel(p.grad, go), n))
    errs.sort()
    median, p90, worst = errs[len(errs) // 2][0], errs[int(len(errs) * 0.9)][0], errs[-1]
    print("%s gradients: %d parameters, l2-rel median %.3g, p90 %.3g, worst %.3g (%s)" % (case, len(errs), median, p90, worst[0], worst[1]))
    assert len(errs) > 550 and median < 3e-2 and p90 < 8e-2 and worst[0] < 0.5, (median, p90, worst)
