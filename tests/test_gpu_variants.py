"""-m gpu: the SURVEY.md 8f-3 variants of the 2-level model through the C ABI - embedding_type 'reduce' (FFHQ checkpoint),
position_embedding '2d', unconditional sos, model_type 'top2bot' (five sequential depth passes) and 'bidirectional' (one
five-token pass) - against goldens made by the unmodified reference and against the oracle."""
from dataclasses import replace

import numpy as np
import pytest
import torch

from oracle import hq_oracle as O
from tests.helpers import build_model, cfg_from_meta, load_golden
from tests.test_oracle_golden import VARIANT_GOLDENS

pytestmark = pytest.mark.gpu

GREEDY = dict(top_k_top=1, top_p_top=1.0, top_k_bot=1, top_p_bot=1.0, softmax_temperature=[1.0, 1.0])


@pytest.mark.parametrize("name", VARIANT_GOLDENS)
@pytest.mark.parametrize("graph,pdl", [(False, False), (True, True)])
def test_variant_greedy_codes_bit_exact_vs_reference_fp32(name, graph, pdl):
    import hqtransformer_b200 as H
    g, meta = load_golden(name)
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="fp32", use_cuda_graph=graph, use_pdl=pdl)
    B = g["codes_top"].shape[0]
    cond = torch.from_numpy(g["labels"]) if cfg.cond == "cls" else None
    ct, cb = H.sampling_ihqgpt(model, B, cond, use_fp16=False, max_seq_len=64, is_tqdm=False, **GREEDY)
    assert np.array_equal(ct.cpu().numpy(), g["codes_top"])
    assert np.array_equal(cb.cpu().numpy(), g["codes_bot"])


@pytest.mark.parametrize("kw,B", [(dict(model_type="top2bot"), 5), (dict(model_type="bidirectional"), 5),
                                  (dict(embedding_type="reduce", cond="uncond"), 4), (dict(position_embedding="2d"), 6),
                                  (dict(model_type="bidirectional", embedding_type="reduce", position_embedding="2d"), 150),
                                  (dict(model_type="top2bot"), 140)])
def test_variant_step_logits_vs_oracle(kw, B):
    """Teacher-forced head outputs: fp32 engine vs the fp32 oracle (< 2e-5), bf16 engine vs the rounding-emulating oracle
    (max-abs <= 2e-2, mean <= 2e-3) - the bars of the 'parallel' model; B > 128 takes the CTA-pair GEMM kernels."""
    import hqtransformer_b200 as H
    cfg = replace(O.SMALL, **kw)
    P = O.make_params(cfg, seed=9, init="rich")
    S = 4
    g = torch.Generator().manual_seed(1)
    cond = torch.randint(0, cfg.n_classes, (B,), generator=g) if cfg.cond == "cls" else None
    ct = torch.randint(0, cfg.vocab_top, (B, S), generator=g)
    cb = torch.randint(0, cfg.vocab_bot, (B, S, 4), generator=g)
    if B <= 8:
        want = O.step_logits(P, cfg, cond, ct, cb)
        m32 = build_model(cfg, P, precision="fp32", max_batch=B, max_seq_len=S)
        got = H.step_logits(m32, cond, ct, cb, use_fp16=False).cpu()
        assert float((got - want).abs().max()) < 2e-5
    emu = O.step_logits(P, cfg, cond, ct, cb, emulate="bf16")
    m16 = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S)
    lg = H.step_logits(m16, cond, ct, cb, use_fp16=True).cpu()
    err = (lg - emu).abs()
    assert float(err.max()) <= 2e-2 and float(err.mean()) <= 2e-3, (float(err.max()), float(err.mean()))


@pytest.mark.parametrize("kw", [dict(model_type="top2bot"), dict(model_type="bidirectional")])
def test_variant_stochastic_sampling_is_deterministic_and_sharding_invariant(kw):
    """bf16, graph + PDL: repeated runs are identical, and a row's codes do not depend on the batch it sits in."""
    import hqtransformer_b200 as H
    cfg = replace(O.SMALL, **kw)
    P = O.make_params(cfg, seed=5, init="rich")
    B = 150
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, cfg.n_classes, (B,), generator=g).cuda()
    model = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=8)
    skw = dict(top_k_top=50, top_p_top=0.9, top_k_bot=50, top_p_bot=0.9, softmax_temperature=[0.9, 0.9], max_seq_len=8,
               is_tqdm=False, use_fp16=True, seed=3)
    ct, cb = H.sampling_ihqgpt(model, B, labels, **skw)
    ct2, cb2 = H.sampling_ihqgpt(model, B, labels, **skw)
    assert torch.equal(ct, ct2) and torch.equal(cb, cb2)
    assert int(ct.max()) < cfg.vocab_top and int(cb.max()) < cfg.vocab_bot and int(ct.min()) >= 0
    cts, cbs = H.sampling_ihqgpt(model, 20, labels[130:150], row_offset=130, **skw)
    assert torch.equal(cts, ct[130:150]) and torch.equal(cbs, cb[130:150])


def test_ffhq_checkpoint_architecture_loads_and_samples():
    """The FFHQ checkpoint's architecture (configs/ffhq_l24.yaml: unconditional, 'reduce' embedding, D = 1024, 16 heads,
    24 + 4 layers): strict parameter set, random init, a short stochastic run."""
    import os
    import hqtransformer_b200 as H
    path = os.path.join(os.path.dirname(H.__file__), "configs", "ffhq_l24.yaml")
    model = H.ImageGPT2.from_config(path, device=0, precision="bf16", max_batch=8).eval()
    s2 = model.stage2
    shapes = s2.param_shapes()
    assert shapes["tok_emb_bot.weight"] == (8192, 256) and "pos_emb_emb.weight" not in shapes and shapes["sos"] == (1, 1, 1024)
    ct, cb = H.sampling_ihqgpt(s2, 8, None, top_k_top=4096, top_k_bot=4096, max_seq_len=4, is_tqdm=False, seed=1)
    assert tuple(ct.shape) == (8, 4) and tuple(cb.shape) == (8, 4, 4) and int(ct.max()) < 8192 and int(ct.min()) >= 0
