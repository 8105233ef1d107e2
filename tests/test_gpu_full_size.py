"""-m gpu: parity at BASELINE's full size (ImageNet L12: D=1536, 24 heads, 12+4 layers, V=8192, 530 M parameters).

The oracle is too slow for a full 256 x 64-position run, so the full-size bars are:
  * the oracle itself on a bounded sample (2 images, 2 positions) against the fp32 and bf16 engines;
  * size-independent properties at B = 256: determinism, graph replay == stream launches, a row's logits do not depend
    on the batch it is computed in (different GEMM tiles / split-K plans), bf16 vs fp32 engine within the bf16 bar.
"""
import pytest
import torch

from oracle import hq_oracle as O
from tests.helpers import build_model

pytestmark = pytest.mark.gpu

CFG = O.IMAGENET_L12


@pytest.fixture(scope="module")
def l12_params():
    return O.make_params(CFG, seed=0, init="reference")


@pytest.fixture(scope="module")
def l12_bf16(l12_params):
    return build_model(CFG, l12_params, precision="bf16", max_batch=256)


def _teacher(B, S, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(0, CFG.n_classes, (B,), generator=g), torch.randint(0, CFG.vocab_top, (B, S), generator=g),
            torch.randint(0, CFG.vocab_bot, (B, S, 4), generator=g))


def test_l12_oracle_sample_vs_fp32_and_bf16_engines(l12_params, l12_bf16):
    import hqtransformer_b200 as H
    labels, ct, cb = _teacher(2, 2, 1)
    ref = O.step_logits(l12_params, CFG, labels, ct, cb)                       # reference algorithm, fp32, CPU
    m32 = build_model(CFG, l12_params, precision="fp32", max_batch=2, max_seq_len=2)
    lg32 = H.step_logits(m32, labels, ct, cb, use_fp16=False).cpu()
    e32 = (lg32 - ref).abs().max().item()
    assert e32 < 5e-5, e32                                                     # fp32 bar (logit std ~0.8)
    assert torch.equal(lg32.argmax(-1), ref.argmax(-1))
    lg16 = H.step_logits(l12_bf16, labels, ct, cb, use_fp16=True).cpu()
    d = (lg16 - ref).abs()
    emu = O.step_logits(l12_params, CFG, labels, ct, cb, emulate="bf16")
    de = (lg16 - emu).abs()
    print(f"L12: fp32 engine vs oracle {e32:.2e}; bf16 engine vs oracle max {d.max():.3e} mean {d.mean():.3e} "
          f"(rel {d.max() / ref.abs().max():.3e}); vs bf16-emulating oracle max {de.max():.3e} mean {de.mean():.3e}")
    assert d.max() <= 0.15 and d.mean() <= 0.02
    assert de.max() <= 5e-2 and de.mean() <= 5e-3


def test_l12_b256_properties(l12_params, l12_bf16):
    import hqtransformer_b200 as H
    B = 256
    labels, ct, cb = _teacher(B, 3, 2)
    kw = dict(top_k_top=2048, top_p_top=0.95, top_k_bot=2048, top_p_bot=0.95, softmax_temperature=[0.95, 0.95],
              use_fp16=True, max_seq_len=64, is_tqdm=False, seed=11)
    a_t, a_b = H.sampling_ihqgpt(l12_bf16, B, labels, **kw)
    b_t, b_b = H.sampling_ihqgpt(l12_bf16, B, labels, **kw)
    assert torch.equal(a_t, b_t) and torch.equal(a_b, b_b)                     # deterministic under a fixed seed
    assert int(a_t.min()) >= 0 and int(a_t.max()) < CFG.vocab_top and int(a_b.max()) < CFG.vocab_bot
    assert a_t.float().std() > 100                                             # not degenerate
    plain = build_model(CFG, l12_params, precision="bf16", max_batch=B, use_cuda_graph=False, use_pdl=False)
    c_t, c_b = H.sampling_ihqgpt(plain, B, labels, **kw)
    assert torch.equal(a_t, c_t) and torch.equal(a_b, c_b)                     # graph + PDL replay == stream launches
    # a row's logits do not depend on the batch it sits in (M = 256 pair kernels + split-K vs M = 40 single-CTA kernels)
    lg_full = H.step_logits(l12_bf16, labels, ct, cb, use_fp16=True)
    lg_part = H.step_logits(l12_bf16, labels[:40], ct[:40], cb[:40], use_fp16=True)
    d = (lg_full[:40] - lg_part).abs()
    print(f"L12 B=256 vs B=40 rows: max {d.max():.3e} mean {d.mean():.3e}")
    assert d.max() <= 5e-2 and d.mean() <= 5e-3
    del plain
    # bf16 engine vs fp32 engine at full size, 16 rows
    m32 = build_model(CFG, l12_params, precision="fp32", max_batch=16, max_seq_len=3)
    lg32 = H.step_logits(m32, labels[:16], ct[:16], cb[:16], use_fp16=False)
    d2 = (lg_full[:16] - lg32).abs()
    print(f"L12 bf16 vs fp32 engine: max {d2.max():.3e} mean {d2.mean():.3e}")
    assert d2.max() <= 0.15 and d2.mean() <= 0.02


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 1 at the reference's real scale: reference-made golden (oracle/make_golden.py --full-size)
# ---------------------------------------------------------------------------------------------------------------------
def test_l12_b4_greedy_grid_bit_exact_vs_reference_golden(l12_params, l12_bf16):
    """ImageNet L12, batch 4 with per-row classes, greedy, all 64 positions: the code grids the UNMODIFIED reference
    produced on CPU (fp32) must be reproduced bit for bit by the fp32 engine; the bf16 engine's teacher-forced logits at
    the stored position must sit within the bf16 bar of the reference's own head outputs."""
    import hqtransformer_b200 as H
    from tests.helpers import load_golden
    g, meta = load_golden("l12_cls_greedy_b4.npz")
    from tests.helpers import cfg_from_meta
    assert cfg_from_meta(meta) == CFG and meta["seed"] == 0 and meta["min_logit_margin"] >= 1e-4
    labels = torch.from_numpy(g["labels"])
    m32 = build_model(CFG, l12_params, precision="fp32", max_batch=4)
    ct, cb = H.sampling_ihqgpt(m32, 4, labels, top_k_top=1, top_p_top=1.0, top_k_bot=1, top_p_bot=1.0, use_fp16=False,
                               max_seq_len=64, is_tqdm=False)
    assert torch.equal(ct.cpu(), torch.from_numpy(g["codes_top"])), "fp32 engine: top grid differs from the reference"
    assert torch.equal(cb.cpu(), torch.from_numpy(g["codes_bot"])), "fp32 engine: bottom grid differs from the reference"
    want = torch.from_numpy(g["logits"])                                       # [4, P, 5, V] reference head outputs
    gct, gcb = torch.from_numpy(g["codes_top"]), torch.from_numpy(g["codes_bot"])
    lg32 = H.step_logits(m32, labels, gct, gcb, use_fp16=False).cpu()[:, meta["logit_positions"]]
    assert (lg32 - want).abs().max() < 5e-5
    del m32
    lg16 = H.step_logits(l12_bf16, labels, gct, gcb, use_fp16=True).cpu()[:, meta["logit_positions"]]
    d = (lg16 - want).abs()
    print(f"L12 golden: bf16 engine vs reference fp32 logits at position {meta['logit_positions']}: max {d.max():.3e} "
          f"mean {d.mean():.3e}")
    assert d.max() <= 0.15 and d.mean() <= 0.02


# ---------------------------------------------------------------------------------------------------------------------
# configs 4 and 5 at full size: bounded-sample parity against the oracle
# ---------------------------------------------------------------------------------------------------------------------
def _bounded_parity(cfg, cond, ct, cb, positions, tag):
    """fp32 engine: identical argmax + fp32 bar; bf16 engine: the bf16 bars (vs fp32 oracle and vs the emulating oracle),
    teacher-forced on `ct` / `cb`, compared at `positions`."""
    import hqtransformer_b200 as H
    P = O.make_params(cfg, seed=1, init="reference")
    B, S = ct.shape
    ref = O.step_logits(P, cfg, cond, ct, cb)[:, positions]
    emu = O.step_logits(P, cfg, cond, ct, cb, emulate="bf16")[:, positions]
    m32 = build_model(cfg, P, precision="fp32", max_batch=B, max_seq_len=S)
    lg32 = H.step_logits(m32, cond, ct, cb, use_fp16=False).cpu()[:, positions]
    e32 = (lg32 - ref).abs().max().item()
    assert e32 < 5e-5, e32
    assert torch.equal(lg32.argmax(-1), ref.argmax(-1))
    del m32
    torch.cuda.empty_cache()
    m16 = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S)
    lg16 = H.step_logits(m16, cond, ct, cb, use_fp16=True).cpu()[:, positions]
    d, de = (lg16 - ref).abs(), (lg16 - emu).abs()
    print(f"{tag}: fp32 engine vs oracle {e32:.2e}; bf16 engine vs oracle max {d.max():.3e} mean {d.mean():.3e}; "
          f"vs bf16-emulating oracle max {de.max():.3e} mean {de.mean():.3e}")
    assert d.max() <= 0.15 and d.mean() <= 0.02
    assert de.max() <= 5e-2 and de.mean() <= 5e-3
    return m16, P


def test_l42_bounded_parity_and_topk_topp_sampling():
    """Config 4: the largest shipped config (L = 42 spatial layers, Ld = 6 depth layers - the 6-fold LayerNorm
    instantiation at full width), 2 images x 2 positions against the oracle, then top-k 2048 / top-p 0.95 / T 0.95
    sampling: deterministic, in range, launch-mode invariant."""
    import hqtransformer_b200 as H
    cfg = O.IMAGENET_L42
    g = torch.Generator().manual_seed(4)
    labels = torch.randint(0, cfg.n_classes, (2,), generator=g)
    ct = torch.randint(0, cfg.vocab_top, (2, 2), generator=g)
    cb = torch.randint(0, cfg.vocab_bot, (2, 2, 4), generator=g)
    m16, P = _bounded_parity(cfg, labels, ct, cb, [0, 1], "L42")
    kw = dict(top_k_top=2048, top_p_top=0.95, top_k_bot=2048, top_p_bot=0.95, softmax_temperature=[0.95, 0.95],
              use_fp16=True, max_seq_len=2, is_tqdm=False, seed=3)
    lab = torch.arange(40) % cfg.n_classes
    a = H.sampling_ihqgpt(m16, 40, lab, **kw)
    b = H.sampling_ihqgpt(m16, 40, lab, **kw)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert 0 <= int(a[0].min()) and int(a[0].max()) < cfg.vocab_top and int(a[1].max()) < cfg.vocab_bot
    del m16
    torch.cuda.empty_cache()
    plain = build_model(cfg, P, precision="bf16", max_batch=40, max_seq_len=2, use_cuda_graph=False, use_pdl=False)
    c = H.sampling_ihqgpt(plain, 40, lab, **kw)
    assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])


def test_cc15m_text_bounded_parity_prefill_and_last_position():
    """Config 5 at full size (CC-15M L12: 64-token text prefix, D = 1536): 2 prompts, teacher-forced over all 64
    positions; compared at the prefill position and at the last decode position (127 cached keys)."""
    cfg = O.CC15M_L12
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, cfg.vocab_txt, (2, cfg.ctx_len_txt), generator=g)
    ct = torch.randint(0, cfg.vocab_top, (2, 64), generator=g)
    cb = torch.randint(0, cfg.vocab_bot, (2, 64, 4), generator=g)
    _bounded_parity(cfg, ids, ct, cb, [0, 1, 63], "CC15M-L12 text")
