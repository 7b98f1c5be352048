"""Host-side logic that needs no GPU: C-ABI surface, module tree / state_dict compatibility with
the reference, optimizer grouping, index maps, ITC queue ring buffer, loud failure without CUDA."""
import os
import re

import pytest
import torch

from oracle import fiber_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _cfg(tasks, image_size=224, L=40):
    import bench
    return bench.config(tasks, image_size, L)


def test_capi_exports_every_declared_symbol():
    from fiber_b200 import lib
    handle = lib.load()  # no GPU needed to load
    syms = lib.declared_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(handle, s), "libfiber_b200.so does not export %s" % s
    assert handle.fiber_version() >= 100


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "fiber_b200.h")).read()
    assert len(re.findall(r"(swin_transformer|roberta|fiber_module)\.py:\d+", text)) >= 8


def test_ctypes_structs_match_header_field_order():
    from fiber_b200 import lib
    text = open(os.path.join(ROOT, "include", "fiber_b200.h")).read()
    for cname, cls in (("fiber_gemm_args", lib.GemmArgs), ("fiber_attn_args", lib.AttnArgs), ("fiber_ln_args", lib.LnArgs),
               ("fiber_ce_args", lib.CeArgs), ("fiber_image_desc", lib.ImageDesc)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"[\*\s]", " ", part).split()[-1])
        assert names == [f[0] for f in cls._fields_], cname


def test_ctypes_struct_layout_matches_the_c_compiler(tmp_path):
    """sizeof / offsetof of every args struct as gcc lays it out == the ctypes mirrors in fiber_b200/lib.py."""
    import ctypes
    import shutil
    import subprocess
    from fiber_b200 import lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = (("fiber_gemm_args", lib.GemmArgs), ("fiber_attn_args", lib.AttnArgs), ("fiber_ln_args", lib.LnArgs),
               ("fiber_ce_args", lib.CeArgs), ("fiber_image_desc", lib.ImageDesc))
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fiber_b200.h"', 'int main(void) {']
    for cname, cls in structs:
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs:
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, "%s.%s" % (cname, fname)


@pytest.mark.parametrize("fname,tasks,size", [("model_cfg0_224_itm_mlm.pt", ["itm", "mlm"], 224),
                                              ("model_384_infer.pt", ["itm", "mlm", "itc"], 384),
                                              ("model_224_vqa.pt", ["vqa"], 224)])
def test_state_dict_matches_reference(fname, tasks, size):
    from fiber_b200.modules import FIBERTransformerSS
    gold = torch.load(os.path.join(GOLD, fname), weights_only=False)
    ref = {k: v for k, v in gold["state_keys"].items() if not re.match(r"(train|val)_", k)}  # PL metric states
    L = gold["L"]
    ours = {k: tuple(v.shape) for k, v in FIBERTransformerSS(_cfg(tasks, size, L)).state_dict().items()}
    assert sorted(ours) == sorted(ref)
    assert all(ours[k] == tuple(ref[k]) for k in ref)


@pytest.mark.parametrize("tag,tasks,size", [("cfg0", ["itm", "mlm"], 224), ("cfg1", ["itm", "mlm", "itc"], 384),
                                            ("vqa", ["vqa"], 224)])
def test_optimizer_groups_match_reference(tag, tasks, size):
    from fiber_b200.modules import FIBERTransformerSS, fiber_utils
    gold = torch.load(os.path.join(GOLD, "schedule.pt"), weights_only=False)[tag]
    model = FIBERTransformerSS(dict(_cfg(tasks, size), learning_rate=1e-5))
    names = {id(p): n for n, p in model.named_parameters()}
    groups = fiber_utils.param_groups(model)
    assert len(groups) == len(gold)
    for g, r in zip(groups, gold):
        assert len(g["params"]) == r["n"] and sum(p.numel() for p in g["params"]) == r["numel"]
        assert abs(g["lr"] - r["lr"]) < 1e-12 and g["weight_decay"] == r["weight_decay"]
        assert sorted(names[id(p)] for p in g["params"])[:3] == r["first"]
    opt, sched = fiber_utils.set_schedule(model)
    assert isinstance(opt[0], torch.optim.AdamW) and sched[0]["interval"] == "step"


def test_module_buffers_match_oracle_index_maps():
    from fiber_b200.modules import swin_transformer as S
    blk = S.SwinTransformerBlock(64, (24, 24), 2, window_size=12, shift_size=6)
    assert torch.equal(blk.attn_mask, O.shift_attn_mask(24, 24, 12, 6))
    assert torch.equal(blk.attn.relative_position_index, O.relative_position_index(12))
    one = S.SwinTransformerBlock(64, (12, 12), 2, window_size=12, shift_size=6)
    assert one.shift_size == 0 and one.attn_mask is None  # swin_transformer.py:304-307


def test_structure_follows_reference_rules():
    from fiber_b200.modules import FIBERTransformerSS
    m = FIBERTransformerSS(_cfg(["itm", "mlm"], 384))
    st2 = m.vit_model.layers[2].blocks
    assert [b.attn.has_i2t for b in st2] == [False] * 14 + [True] * 4
    assert [b.shift_size for b in st2[14:]] == [0, 6, 0, 6]
    assert all(b.attn.has_i2t and b.shift_size == 0 for b in m.vit_model.layers[3].blocks)
    enc = m.text_transformer.encoder.layer
    assert [hasattr(l, "crossattention_t2i") for l in enc] == [False] * 6 + [True] * 6
    assert enc[6].crossattention_t2i.self.key.weight.shape == (768, 512)
    assert enc[10].crossattention_t2i.self.key.weight.shape == (768, 1024)
    # rank_output aliases row 1 of the ITM classifier (fiber_module.py:112-114)
    assert m.rank_output.weight.data_ptr() == m.itm_score.fc.weight[1:].data_ptr()
    dpr = [b.drop_path.drop_prob if hasattr(b.drop_path, "drop_prob") else 0.0
           for layer in m.vit_model.layers for b in layer.blocks]
    assert dpr[0] == 0.0 and abs(dpr[-1] - 0.1) < 1e-6 and dpr == sorted(dpr)


def test_extended_mask_and_position_ids_semantics():
    from fiber_b200.modules.roberta import RobertaConfig, RobertaModel
    cfg = RobertaConfig(num_hidden_layers=1, vocab_size=50)
    rm = RobertaModel(cfg, add_pooling_layer=False)
    mask = torch.tensor([[1, 1, 0], [1, 0, 0]])
    ext = rm.get_extended_attention_mask(mask, mask.shape, mask.device)
    assert ext.shape == (2, 1, 1, 3) and torch.equal(ext, O.extended_mask(mask))
    assert float(ext.min()) == -10000.0 and float(ext.max()) == 0.0


def test_itc_queue_ring_buffer_wraps():
    from fiber_b200.modules import FIBERTransformerSS
    m = FIBERTransformerSS(_cfg(["itc"], 32))
    m.queue_size = 10
    for name in ("image_queue", "text_queue"):
        setattr(m, name, torch.zeros(768, 10))
    m.image_input_queue = torch.zeros(10, 3, 32, 32)
    m.text_input_queue = torch.zeros(10, 40, dtype=torch.long)
    m.text_input_mask_queue = torch.zeros(10, 40, dtype=torch.long)
    for step in range(3):
        n = 4
        f = torch.full((n, 768), float(step + 1))
        m._dequeue_and_enqueue(f, -f, torch.full((n, 3, 32, 32), float(step + 1)),
                               torch.full((n, 40), step + 1), torch.ones(n, 40, dtype=torch.long))
    assert int(m.queue_ptr) == 2 and int(m.queue_total) == 12
    assert m.image_queue[0].tolist() == [3, 3, 1, 1, 2, 2, 2, 2, 3, 3]
    assert m.text_input_queue[:, 0].tolist() == [3, 3, 1, 1, 2, 2, 2, 2, 3, 3]


def test_no_cpu_fallback():
    """The product path must fail loudly instead of computing on the CPU."""
    from fiber_b200.modules import swin_transformer as S
    blk = S.SwinTransformerBlock(64, (14, 14), 2, window_size=7)
    with pytest.raises(RuntimeError, match="CUDA"):
        blk(torch.randn(1, 196, 64, dtype=torch.bfloat16))


def test_product_code_never_imports_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "fiber_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(d, f)).read(), os.path.join(d, f)


def test_rows_of_cat_equals_cat_then_index():
    """objectives._rows_of_cat == torch.cat([batch, queue])[idx] (objectives.py:142-166 of the reference), including
    the empty-queue case, without materialising the concatenation."""
    import torch
    from fiber_b200.modules.objectives import _rows_of_cat
    g = torch.Generator().manual_seed(0)
    for shape in ((5, 3, 4, 4), (5, 7)):
        for qn in (0, 1, 9):
            if len(shape) == 4:
                a, b = torch.randn(shape, generator=g), torch.randn((qn,) + shape[1:], generator=g)
            else:
                a = torch.randint(0, 100, shape, generator=g)
                b = torch.randint(0, 100, (qn,) + shape[1:], generator=g)
            idx = torch.randint(0, shape[0] + qn, (shape[0],), generator=g)
            assert torch.equal(_rows_of_cat(a, b, idx), torch.cat([a, b], 0)[idx])


# ---------------------------------------------------------------------------------------------
# Round 2: checkpoint adaptation, schedules, epoch bookkeeping, pretrained-weight loading
# ---------------------------------------------------------------------------------------------
def test_load_adapts_384_checkpoint_to_576(tmp_path):
    """task_finetune_vqa load_path=<384-px pre-training ckpt> at image_size=576 (fiber_module.py:139-148,
    swin_helpers.py:20-44): relative-position tables are resized bicubically, masks / index buffers rebuilt."""
    from fiber_b200.modules import FIBERTransformerSS
    from fiber_b200.modules.swin_transformer import swin_adapt_position_encoding
    src = FIBERTransformerSS(_cfg(["itm", "mlm", "itc"], 384))
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    ck = tmp_path / "pretrain384.ckpt"
    torch.save({"state_dict": sd}, ck)
    cfg = dict(_cfg(["vqa"], 576, 50), load_path=str(ck), resolution_before=384)
    dst = FIBERTransformerSS(cfg)  # would raise size-mismatch errors without the adaptation
    key = "vit_model.layers.2.blocks.3.attn.relative_position_bias_table"
    assert sd[key].shape == (23 * 23, 16) and dst.state_dict()[key].shape == (35 * 35, 16)
    # same arithmetic as the reference's helper: bicubic resize of the (heads, 23, 23) grid
    want = torch.nn.functional.interpolate(sd[key].t().reshape(1, 16, 23, 23), size=(35, 35), mode="bicubic")[0]
    torch.testing.assert_close(dst.state_dict()[key], want.permute(1, 2, 0).reshape(35 * 35, 16))
    # non-resolution parameters arrive unchanged; the 576-px masks were rebuilt, not loaded
    k2 = "text_transformer.encoder.layer.7.crossattention_t2i.self.key.weight"
    assert torch.equal(dst.state_dict()[k2], sd[k2])
    assert dst.state_dict()["vit_model.layers.0.blocks.1.attn_mask"].shape == (64, 324, 324)
    # same resolution: identity
    same = {"a.relative_position_bias_table": torch.randn(529, 4)}
    assert swin_adapt_position_encoding(dict(same), before=384, after=384)["a.relative_position_bias_table"] is \
        same["a.relative_position_bias_table"]


def test_set_schedule_derives_max_steps_and_supports_cosine():
    """Every fine-tuning config has max_steps=None and warmup_steps=0.1 (config.py:134-150): max_steps comes from the
    trainer's dataloader length (fiber_utils.py:254-262); decay_power='cosine' selects the cosine schedule."""
    import math
    import types
    from fiber_b200.modules import FIBERTransformerSS, fiber_utils
    cfg = dict(_cfg(["vqa"], 224, 50), max_steps=None, warmup_steps=0.1, decay_power="cosine")
    model = FIBERTransformerSS(cfg)
    with pytest.raises(ValueError):
        fiber_utils.set_schedule(model)  # no trainer to derive max_steps from: loud, not a TypeError
    dm = types.SimpleNamespace(train_dataloader=lambda: range(250))
    model.trainer = types.SimpleNamespace(max_steps=None, max_epochs=8, accumulate_grad_batches=2, datamodule=dm)
    (opt,), (sched,) = fiber_utils.set_schedule(model)
    assert fiber_utils.resolve_max_steps(model) == 250 * 8 // 2 == 1000
    lam = sched["scheduler"].lr_lambdas[0]
    assert lam(0) == 0.0 and lam(50) == pytest.approx(0.5) and lam(100) == pytest.approx(1.0)
    assert lam(550) == pytest.approx(0.5 * (1 + math.cos(math.pi * 0.5))) and lam(1000) == pytest.approx(0.0, abs=1e-12)
    # polynomial branch with trainer.max_steps given (pre-training configs)
    model.hparams.config["decay_power"] = 1
    model.trainer.max_steps = 200
    (_,), (sched,) = fiber_utils.set_schedule(model)
    lam = sched["scheduler"].lr_lambdas[0]
    assert lam(10) == pytest.approx(0.5) and lam(110) == pytest.approx(0.5) and lam(200) == pytest.approx(0.0)


def test_metrics_log_batch_values_and_epoch_wrapup_resets():
    """PL Metric.forward returns the value of the CURRENT batch and accumulates the epoch state; epoch_wrapup logs the
    epoch values, resets them and logs '<phase>/the_metric' (what run.py's ModelCheckpoint monitors)."""
    from fiber_b200.modules import FIBERTransformerSS, fiber_utils
    from fiber_b200.modules.lightning import Accuracy, Scalar
    acc = Accuracy()
    logits = torch.tensor([[2.0, 1.0], [0.0, 1.0], [3.0, 0.0], [0.0, 5.0]])
    assert float(acc(logits, torch.tensor([0, 1, 1, -100]))) == pytest.approx(2 / 3)   # batch 1: 2 of 3 counted
    assert float(acc(logits, torch.tensor([1, 0, 1, 0]))) == pytest.approx(0.0)        # batch 2 alone, not the running mean
    assert float(acc.compute()) == pytest.approx(2 / 7)
    sc = Scalar()
    assert float(sc(torch.tensor(4.0))) == 4.0 and float(sc(2.0)) == 2.0 and float(sc.compute()) == 3.0
    model = FIBERTransformerSS(_cfg(["itm", "mlm", "itc"]))
    model.train()
    model.train_itm_accuracy(logits, torch.tensor([0, 1, 0, 1]))
    model.train_mlm_accuracy(logits, torch.tensor([0, 0, 0, 0]))
    model.train_itc_t2i_accuracy(logits, torch.tensor([0, 1, 0, 0]))
    for n in ("itm", "mlm", "itc"):
        getattr(model, "train_%s_loss" % n)(torch.tensor(1.5))
    model.training_epoch_end([])
    assert float(model.logged["train/the_metric"]) == pytest.approx(1.0 + 0.5 + 0.75)
    assert float(model.logged["itm/train/loss_epoch"]) == 1.5 and "itc/train/t2i_accuracy_epoch" in model.logged
    assert float(model.train_itm_accuracy.total) == 0.0  # reset
    model.eval()
    model.val_itm_accuracy(logits, torch.tensor([0, 1, 0, 1]))
    model.validation_epoch_end([])
    assert "val/the_metric" in model.logged
    assert model.validation_step.__doc__ is None or True


def test_from_pretrained_warns_loudly_and_loads_local_weights(tmp_path, monkeypatch):
    from fiber_b200.modules.roberta import RobertaModel
    monkeypatch.delenv("FIBER_ROBERTA_WEIGHTS", raising=False)
    with pytest.warns(RuntimeWarning, match="RANDOMLY INITIALISED"):
        ref = RobertaModel.from_pretrained("roberta-base")
    with pytest.raises(ValueError):
        RobertaModel.from_pretrained("bert-base-uncased")
    # an HF-style checkpoint (keys prefixed "roberta.", an lm_head, no t2i tensors) on local disk
    sd = {"roberta." + k: torch.full_like(v, 0.25) for k, v in ref.state_dict().items() if "t2i" not in k}
    sd["lm_head.bias"] = torch.zeros(3)
    path = tmp_path / "pytorch_model.bin"
    torch.save(sd, path)
    monkeypatch.setenv("FIBER_ROBERTA_WEIGHTS", str(path))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        m = RobertaModel.from_pretrained("roberta-base")
    assert float(m.encoder.layer[3].intermediate.dense.weight.mean()) == 0.25
    assert float(m.encoder.layer[7].crossattention_t2i.self.key.weight.abs().max()) != 0.25  # t2i stays freshly initialised


def test_dropout_seed_follows_torch_seed():
    from fiber_b200 import ops
    ops._base_seed = None
    torch.manual_seed(123)
    a = ops._resolve_base_seed()
    torch.manual_seed(124)
    b = ops._resolve_base_seed()
    torch.manual_seed(123)
    assert ops._resolve_base_seed() == a != b
    ops.set_dropout_seed(7)
    assert ops._next_seed() == (7 * 1000003 + 1) & 0x7FFFFFFFFFFF
    ops._base_seed = None


def test_fused_adamw_tables_cover_every_element():
    """Host half of the fused multi-tensor AdamW (fiber_b200/optim.py): table layout == the header's struct, chunks tile
    every tensor exactly once."""
    import numpy as np
    from fiber_b200 import optim
    text = open(os.path.join(ROOT, "include", "fiber_b200.h")).read()
    body = re.search(r"typedef struct fiber_adamw_tensor \{(.*?)\} fiber_adamw_tensor;", text, re.S).group(1)
    assert re.findall(r"(\w+)\s*[;,]", body) == ["p", "g", "m", "v", "p_bf16", "n", "lr", "wd"]
    sizes = [1, 1023, 65536, 65537, 3 * 65536 + 5]
    t, ch = optim.build_tables((8 * i, 16 * i, 24 * i, 32 * i, 0, n, 1e-5 * (i + 1), 0.01 * (i % 2)) for i, n in enumerate(sizes))
    assert t.dtype.itemsize == 56 and list(t["n"]) == sizes and t["lr"][2] == np.float32(3e-5)
    covered = {i: 0 for i in range(len(sizes))}
    for ti, ci in ch:
        covered[int(ti)] += min(optim.CHUNK, sizes[ti] - ci * optim.CHUNK)
    assert covered == dict(enumerate(sizes)) and len(ch) == 1 + 1 + 1 + 2 + 4
    with pytest.raises(RuntimeError):  # no CPU path
        p = torch.nn.Parameter(torch.zeros(4))
        p.grad = torch.ones(4)
        optim.FusedAdamW([p], lr=1e-3).step()
