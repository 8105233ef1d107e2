"""-m gpu: the whole sampling loop through the drop-in API against the reference-made goldens and the oracle."""
import numpy as np
import pytest
import torch

from oracle import hq_oracle as O
from tests.helpers import build_model, cfg_from_meta, load_golden

pytestmark = pytest.mark.gpu

GREEDY = dict(top_k_top=1, top_p_top=1.0, top_k_bot=1, top_p_bot=1.0, softmax_temperature=[1.0, 1.0])


@pytest.mark.parametrize("name", ["tiny_cls_greedy.npz", "small_cls_greedy.npz", "asym_cls_greedy.npz"])
@pytest.mark.parametrize("graph,pdl", [(False, False), (True, False), (False, True), (True, True)])
def test_greedy_codes_bit_exact_vs_reference_fp32(name, graph, pdl):
    """Config 1 of BASELINE.json: greedy code grids from the reference's own sampler (CPU, fp32) must be reproduced
    bit for bit by the fp32 engine (use_fp16=False).  Golden margins are >= 1e-4, fp32 GEMM error ~1e-6."""
    import hqtransformer_b200 as H
    g, meta = load_golden(name)
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="fp32", use_cuda_graph=graph, use_pdl=pdl)
    labels = torch.from_numpy(g["labels"])
    ct, cb = H.sampling_ihqgpt(model, len(labels), labels, use_fp16=False, max_seq_len=64, is_tqdm=False, **GREEDY)
    assert ct.dtype == torch.int64 and tuple(ct.shape) == (len(labels), 64) and tuple(cb.shape) == (len(labels), 64, 4)
    assert np.array_equal(ct.cpu().numpy(), g["codes_top"])
    assert np.array_equal(cb.cpu().numpy(), g["codes_bot"])
    # the reference's own driver broadcasts ONE class to the batch (sampling.py:183-186)
    cls = int(labels[-1])
    ct_s, cb_s = H.sampling_ihqgpt(model, len(labels), cls, use_fp16=False, max_seq_len=64, is_tqdm=False, **GREEDY)
    assert np.array_equal(ct_s.cpu().numpy(), g["codes_top_scalar_class"])
    assert np.array_equal(cb_s.cpu().numpy(), g["codes_bot_scalar_class"])


@pytest.mark.parametrize("name", ["tiny_cls_greedy.npz", "small_cls_greedy.npz", "asym_cls_greedy.npz"])
def test_step_logits_fp32_vs_reference(name):
    """Per-position head outputs captured from the reference (hooks on head_top/head_bot) at 5 positions."""
    import hqtransformer_b200 as H
    g, meta = load_golden(name)
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="fp32")
    lg = H.step_logits(model, torch.from_numpy(g["labels"]), torch.from_numpy(g["codes_top"]),
                       torch.from_numpy(g["codes_bot"]), use_fp16=False).cpu().numpy()
    want = g["logits"]                                  # [B, P, 5, V]
    got = lg[:, meta["logit_positions"]]
    err = np.abs(got - want).max()
    assert err < 2e-5, err                              # stated fp32 tolerance: 2e-5 absolute (logit scale ~1)


@pytest.mark.parametrize("name", ["tiny_cls_greedy.npz", "small_cls_greedy.npz", "asym_cls_greedy.npz"])
def test_step_logits_bf16_vs_oracle_and_reference(name):
    """bf16 production path (tcgen05 GEMMs, bf16 KV cache).  Two bars, both written here:
    (a) against the oracle run with the SAME rounding points (emulate='bf16'): max-abs <= 2e-2, mean-abs <= 2e-3
        (differences come only from fp32 summation order flipping a bf16 rounding now and then);
    (b) against the reference's fp32 logits: max-abs <= 0.15, mean-abs <= 0.02 at logit std ~0.3-0.8 (the bf16
        tolerance of north_star; fp16 autocast in the reference has the same order of error)."""
    import hqtransformer_b200 as H
    g, meta = load_golden(name)
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="bf16")
    labels, ct, cb = torch.from_numpy(g["labels"]), torch.from_numpy(g["codes_top"]), torch.from_numpy(g["codes_bot"])
    lg = H.step_logits(model, labels, ct, cb, use_fp16=True).cpu()
    emu = O.step_logits(P, cfg, labels, ct, cb, emulate="bf16")
    d = (lg - emu).abs()
    print(f"{name}: bf16 vs emulated oracle max-abs {d.max():.3e} mean-abs {d.mean():.3e}")
    assert d.max() <= 2e-2 and d.mean() <= 2e-3, (d.max(), d.mean())
    want = torch.from_numpy(g["logits"])
    d2 = (lg[:, meta["logit_positions"]] - want).abs()
    rel = d2.max() / want.abs().max()
    print(f"{name}: bf16 vs reference fp32 max-abs {d2.max():.3e} mean-abs {d2.mean():.3e} max-rel {rel:.3e}")
    assert d2.max() <= 0.15 and d2.mean() <= 0.02, (d2.max(), d2.mean())


@pytest.mark.parametrize("graph,pdl", [(False, False), (True, True)])
def test_bf16_sampling_is_deterministic_and_launch_mode_invariant(graph, pdl):
    """Same seed -> same grids, whether the loop is replayed as a CUDA graph with programmatic dependent launch or
    issued as plain stream launches (a race in the PDL chain would show up here)."""
    import hqtransformer_b200 as H
    g, meta = load_golden("small_cls_greedy.npz")
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    labels = torch.arange(300) % cfg.n_classes
    kw = dict(top_k_top=50, top_p_top=0.9, top_k_bot=50, top_p_bot=0.9, softmax_temperature=[0.9, 0.9], use_fp16=True,
              max_seq_len=16, is_tqdm=False, seed=7)
    base = build_model(cfg, P, precision="bf16", max_batch=300, use_cuda_graph=False, use_pdl=False)
    ct0, cb0 = H.sampling_ihqgpt(base, 300, labels, **kw)
    model = build_model(cfg, P, precision="bf16", max_batch=300, use_cuda_graph=graph, use_pdl=pdl)
    for _ in range(3):
        ct, cb = H.sampling_ihqgpt(model, 300, labels, **kw)
        assert torch.equal(ct, ct0) and torch.equal(cb, cb0)


@pytest.mark.parametrize("cuts", [(0, 19, 37), (0, 1, 36, 37), (0, 150, 300)])
def test_bf16_results_do_not_depend_on_how_the_batch_is_cut(cuts):
    """A row's codes are a function of (weights, its conditioning, seed, GLOBAL row index) only: sampling the batch in
    one call equals sampling it shard by shard (what `sampling_ihqgpt_sharded` does across GPUs), bit for bit, greedy
    and stochastic, although the shards run other GEMM kernels (single CTA for <= 128 rows, CTA pairs above).  Holds
    because the split-K factor of the residual GEMMs depends on the weight shape only (engine.cu: resid_splits)."""
    import hqtransformer_b200 as H
    cfg = O.SMALL
    P = O.make_params(cfg, seed=5, init="rich")
    B = cuts[-1]
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, cfg.n_classes, (B,), generator=g).cuda()
    model = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=16)
    for kw in (dict(top_k_top=1, top_k_bot=1),
               dict(top_k_top=50, top_p_top=0.9, top_k_bot=50, top_p_bot=0.9, softmax_temperature=[0.9, 0.9])):
        ct, cb = H.sampling_ihqgpt(model, B, labels, max_seq_len=16, is_tqdm=False, use_fp16=True, seed=3, **kw)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            ct_s, cb_s = H.sampling_ihqgpt(model, hi - lo, labels[lo:hi], max_seq_len=16, is_tqdm=False, use_fp16=True,
                                           seed=3, row_offset=lo, **kw)
            assert torch.equal(ct_s, ct[lo:hi]) and torch.equal(cb_s, cb[lo:hi]), (lo, hi, kw)


@pytest.mark.parametrize("B,force_split", [(40, None), (150, None), (40, "2"), (150, "2")])
def test_step_logits_bf16_wide_batch_vs_oracle(B, force_split, monkeypatch):
    """Batches wide enough to take the CTA-pair GEMM kernel (M > 128) and the split-K fc2 + LayerNorm fold path;
    same tolerance as the narrow-batch logits test."""
    import hqtransformer_b200 as H
    if force_split is not None:
        monkeypatch.setenv("HQ_DEBUG", "1")                    # switches are honoured only in debug mode ...
        monkeypatch.setenv("HQ_FORCE_SPLITK", force_split)      # ... pin the split-K factor of the fc2 GEMMs
    g, meta = load_golden("small_cls_greedy.npz")
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    S = 3
    gen = torch.Generator().manual_seed(B)
    labels = torch.randint(0, cfg.n_classes, (B,), generator=gen)
    ct = torch.randint(0, cfg.vocab_top, (B, S), generator=gen)
    cb = torch.randint(0, cfg.vocab_bot, (B, S, 4), generator=gen)
    model = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S)
    lg = H.step_logits(model, labels, ct, cb, use_fp16=True).cpu()
    emu = O.step_logits(P, cfg, labels, ct, cb, emulate="bf16")
    d = (lg - emu).abs()
    print(f"B={B}: bf16 vs emulated oracle max-abs {d.max():.3e} mean-abs {d.mean():.3e}")
    assert d.max() <= 2e-2 and d.mean() <= 2e-3, (d.max(), d.mean())
    # fp32 engine on the same inputs: plain fp32 tolerance
    m32 = build_model(cfg, P, precision="fp32", max_batch=B, max_seq_len=S)
    lg32 = H.step_logits(m32, labels, ct, cb, use_fp16=False).cpu()
    ref = O.step_logits(P, cfg, labels, ct, cb)
    assert (lg32 - ref).abs().max() < 2e-5


def test_bf16_greedy_margin_aware():
    """bf16 cannot be bit-exact on random weights: every greedy disagreement with the fp32 reference
    must sit at a position where the reference's own top-1/top-2 margin is below the bf16 logit tolerance."""
    import hqtransformer_b200 as H
    g, meta = load_golden("small_cls_greedy.npz")
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="bf16")
    labels, ct, cb = torch.from_numpy(g["labels"]), torch.from_numpy(g["codes_top"]), torch.from_numpy(g["codes_bot"])
    lg = H.step_logits(model, labels, ct, cb, use_fp16=True).cpu()       # teacher-forced on the reference's codes
    ref = O.step_logits(P, cfg, labels, ct, cb)
    mine, theirs = lg.argmax(-1), ref.argmax(-1)
    top2 = ref.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    bad = mine != theirs
    print(f"bf16 greedy agreement {(~bad).float().mean():.4f}; largest margin among flips "
          f"{margin[bad].max().item() if bad.any() else 0:.3e}")
    assert (~bad).float().mean() > 0.9
    assert not bad.any() or margin[bad].max() < 0.15


@pytest.mark.parametrize("name", ["tiny_txt_greedy.npz", "asym_txt_greedy.npz"])
@pytest.mark.parametrize("graph,pdl", [(False, False), (True, False), (False, True), (True, True)])
def test_text_prefix_greedy_bit_exact_fp32(name, graph, pdl):
    """Config 5 shape (text prefix -> 64-token causal prefill of B*64 rows, then cached decode over 64..127 keys) vs the
    reference-made goldens, in every launch mode; the asymmetric model (L != Ld, vocab_top != vocab_bot, 6 heads, B = 2)
    is the one that caught the cross-position PDL race on the class-conditional path."""
    import hqtransformer_b200 as H
    g, meta = load_golden(name)
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="fp32", use_cuda_graph=graph, use_pdl=pdl)
    ids = torch.from_numpy(g["text_ids"])
    for _ in range(2):                                   # a second run on the same ctx must not see the first one's cache
        ct, cb = H.sampling_ihqgpt(model, 1, ids, use_fp16=False, max_seq_len=64, is_tqdm=False, **GREEDY)
        assert np.array_equal(ct.cpu().numpy(), g["codes_top"]) and np.array_equal(cb.cpu().numpy(), g["codes_bot"])


@pytest.mark.parametrize("name", ["tiny_txt_greedy.npz", "asym_txt_greedy.npz"])
def test_text_prefix_bf16_logits_and_launch_mode_invariance(name):
    """bf16 text path: teacher-forced logits against the emulating oracle, and stochastic sampling of a B > 1 batch that
    must not change across launch modes (graph + PDL vs plain stream launches) or repeats."""
    import hqtransformer_b200 as H
    g, meta = load_golden(name)
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    ids = torch.from_numpy(g["text_ids"])
    model16 = build_model(cfg, P, precision="bf16")
    ctg, cbg = torch.from_numpy(g["codes_top"]), torch.from_numpy(g["codes_bot"])
    lg = H.step_logits(model16, ids, ctg, cbg, use_fp16=True).cpu()
    emu = O.step_logits(P, cfg, ids, ctg, cbg, emulate="bf16")
    d = (lg - emu).abs()
    assert d.max() <= 2e-2 and d.mean() <= 2e-3, (d.max(), d.mean())
    B = 6
    gen = torch.Generator().manual_seed(3)
    ids6 = torch.randint(0, cfg.vocab_txt, (B, cfg.ctx_len_txt), generator=gen)
    kw = dict(top_k_top=50, top_p_top=0.9, top_k_bot=50, top_p_bot=0.9, softmax_temperature=[0.9, 0.9], use_fp16=True,
              max_seq_len=64, is_tqdm=False, seed=21)
    plain = build_model(cfg, P, precision="bf16", use_cuda_graph=False, use_pdl=False)
    ct0, cb0 = H.sampling_ihqgpt(plain, B, ids6, **kw)
    for graph, pdl in ((True, True), (True, False), (False, True)):
        m = build_model(cfg, P, precision="bf16", use_cuda_graph=graph, use_pdl=pdl)
        for _ in range(2):
            ct, cb = H.sampling_ihqgpt(m, B, ids6, **kw)
            assert torch.equal(ct, ct0) and torch.equal(cb, cb0), (graph, pdl)


def test_default_seed_advances_like_the_reference_generator():
    """The reference draws with `torch.multinomial` from the global generator: consecutive calls with identical arguments
    give different samples (sampling_hqmodel.py:180-193 relies on it) and `set_seed` reproduces the whole sequence of
    calls.  Here the default Philox key of a call is drawn from torch's default generator."""
    import hqtransformer_b200 as H
    g, meta = load_golden("tiny_uncond_stochastic.npz")
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="fp32", max_batch=64, max_seq_len=4)
    kw = dict(use_fp16=False, max_seq_len=4, is_tqdm=False, top_k_top=20, top_k_bot=20)
    torch.manual_seed(77)
    a = H.sampling_ihqgpt(model, 64, None, **kw)
    b = H.sampling_ihqgpt(model, 64, None, **kw)
    assert not torch.equal(a[0], b[0]), "two default-seed calls returned the same grids"
    torch.manual_seed(77)
    a2 = H.sampling_ihqgpt(model, 64, None, **kw)
    b2 = H.sampling_ihqgpt(model, 64, None, **kw)
    assert torch.equal(a[0], a2[0]) and torch.equal(a[1], a2[1]) and torch.equal(b[0], b2[0]) and torch.equal(b[1], b2[1])
    # the per-position API draws ONE key per batch (at past=None) and keeps it for the rest of the batch
    sos = P["sos"].repeat(8, 1, 1).cuda()
    torch.manual_seed(5)
    t0, b0, past = model.sampling_step(sos, None, None, None, use_fp16=False, top_k_top=20, top_k_bot=20, past=None)
    k0 = model._step_state["seed"]
    t1, b1, past = model.sampling_step(sos, t0, b0[:, 0], torch.zeros(8, 1, dtype=torch.long), use_fp16=False,
                                       top_k_top=20, top_k_bot=20, past=past)
    assert model._step_state["seed"] == k0
    torch.manual_seed(5)
    t0b, _, _ = model.sampling_step(sos, None, None, None, use_fp16=False, top_k_top=20, top_k_bot=20, past=None)
    assert torch.equal(t0, t0b) and model._step_state["seed"] == k0


def test_given_top_code_broadcast_row():
    """hierarchical_ar.py:771-772 accepts a [1, S] given_top_code for a batch of B rows."""
    import hqtransformer_b200 as H
    g, meta = load_golden("tiny_cls_greedy.npz")
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="fp32")
    given = torch.from_numpy(g["codes_top"])[:1]
    ct, _ = H.sampling_ihqgpt(model, 3, 4, use_fp16=False, max_seq_len=64, is_tqdm=False, given_top_code=given, **GREEDY)
    assert torch.equal(ct.cpu(), given.expand(3, -1))


def test_sampling_step_api_matches_full_loop():
    """`iHQGPT.sampling_step` driven position by position with the reference's outer loop (sampling.py:194-234)
    gives the same codes as one `sampling_ihqgpt` call."""
    import hqtransformer_b200 as H
    g, meta = load_golden("tiny_cls_greedy.npz")
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="fp32")
    labels = torch.from_numpy(g["labels"])
    sos = P["sos.weight"][labels].unsqueeze(1).cuda()
    codes_top = codes_bot = past = None
    for cnt in range(64):
        if codes_top is None:
            ct_ = cb_ = pos_ = None
        else:
            ct_ = codes_top[:, cnt - 1:cnt]
            cb_ = codes_bot[:, cnt - 1, :]
            pos_ = H.get_positional_encoding(codes_top, mode="1d")[:, cnt - 1:cnt]
        code_top, code_bot, present = model.sampling_step(sos=sos, codes_t=ct_, codes_b=cb_, pos_codes=pos_,
                                                          use_fp16=False, past=past, **GREEDY)
        past = [present] if past is None else past + [present]
        codes_top = code_top if codes_top is None else torch.cat([codes_top, code_top], 1)
        codes_bot = code_bot if codes_bot is None else torch.cat([codes_bot, code_bot], 1)
    assert np.array_equal(codes_top.cpu().numpy(), g["codes_top"])
    assert np.array_equal(codes_bot.cpu().numpy(), g["codes_bot"])


def test_stochastic_token_statistics_vs_reference_distribution():
    """Stochastic sampling (top-k / top-p / temperature): Philox != torch's generator, so compare statistics.
    At position 0 every row sees the same logits (unconditional model), so the empirical top-code frequencies over a
    large batch must match the reference's filtered distribution (oracle == reference on CPU) - chi-square test; and
    the marginal over the 4 bottom codes is checked the same way given the most frequent top code."""
    import hqtransformer_b200 as H
    g, meta = load_golden("tiny_uncond_stochastic.npz")
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    kw = dict(top_k_top=meta["top_k_top"], top_p_top=meta["top_p_top"], top_k_bot=meta["top_k_bot"],
              top_p_bot=meta["top_p_bot"], softmax_temperature=meta["softmax_temperature"])
    B = 4096
    model = build_model(cfg, P, precision="fp32", max_batch=B, max_seq_len=2)
    ct, cb = H.sampling_ihqgpt(model, B, None, use_fp16=False, max_seq_len=2, is_tqdm=False, seed=99, **kw)
    ct, cb = ct.cpu(), cb.cpu()
    # reference distribution of the first top code
    _, _, lg = O.sample(P, cfg, None, 1, max_seq_len=1, return_logits=True, top_k_top=1, top_k_bot=1)
    _, pr = O.draw_token(lg[:, 0, 0, :cfg.vocab_top], kw["softmax_temperature"][0], kw["top_k_top"], kw["top_p_top"])
    pr = pr[0].double().numpy()
    counts = np.bincount(ct[:, 0].numpy(), minlength=cfg.vocab_top).astype(np.float64)
    assert set(np.nonzero(counts)[0]) <= set(np.nonzero(pr)[0])
    keep = pr * B >= 5
    chi2 = (((counts - pr * B) ** 2)[keep] / (pr * B)[keep]).sum()
    dof = max(int(keep.sum()) - 1, 1)
    assert chi2 < dof + 5 * (2 * dof) ** 0.5, (chi2, dof)
    # determinism under a fixed seed, and a different stream under another seed
    ct2, cb2 = H.sampling_ihqgpt(model, B, None, use_fp16=False, max_seq_len=2, is_tqdm=False, seed=99, **kw)
    assert torch.equal(ct2.cpu(), ct) and torch.equal(cb2.cpu(), cb)
    ct3, _ = H.sampling_ihqgpt(model, B, None, use_fp16=False, max_seq_len=2, is_tqdm=False, seed=100, **kw)
    assert not torch.equal(ct3.cpu(), ct)
    # rows [lo, hi) sampled as a shard with row_offset reproduce the same rows (multi-GPU invariance)
    ct4, cb4 = H.sampling_ihqgpt(model, 100, None, use_fp16=False, max_seq_len=2, is_tqdm=False, seed=99,
                                 row_offset=1000, **kw)
    assert torch.equal(ct4.cpu(), ct[1000:1100]) and torch.equal(cb4.cpu(), cb[1000:1100])


def test_given_top_code_is_respected():
    """given_top_code (sampling.py:205-208, hierarchical_ar.py:771-772): top codes forced, bottoms sampled."""
    import hqtransformer_b200 as H
    g, meta = load_golden("tiny_cls_greedy.npz")
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    model = build_model(cfg, P, precision="fp32")
    labels = torch.from_numpy(g["labels"])
    given = torch.from_numpy(g["codes_top"])
    ct, cb = H.sampling_ihqgpt(model, len(labels), labels, use_fp16=False, max_seq_len=64, is_tqdm=False,
                               given_top_code=given, **GREEDY)
    assert np.array_equal(ct.cpu().numpy(), g["codes_top"]) and np.array_equal(cb.cpu().numpy(), g["codes_bot"])
    given2 = (given + 1) % cfg.vocab_top
    ct2, cb2 = H.sampling_ihqgpt(model, len(labels), labels, use_fp16=False, max_seq_len=64, is_tqdm=False,
                                 given_top_code=given2, **GREEDY)
    assert torch.equal(ct2.cpu(), given2)
    want_t, want_b = O.sample(P, cfg, labels, len(labels), top_k_top=1, top_k_bot=1, given_top_code=given2)
    assert torch.equal(cb2.cpu(), want_b)


def test_errors_are_loud():
    import hqtransformer_b200 as H
    g, meta = load_golden("tiny_cls_greedy.npz")
    cfg = cfg_from_meta(meta)
    P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
    bad = dict(P)
    bad.pop("ln_f.bias")
    with pytest.raises(H.HQError, match="missing"):
        build_model(cfg, bad, precision="fp32")
    bad = dict(P)
    bad["extra.weight"] = torch.zeros(3)
    with pytest.raises(H.HQError, match="unexpected key"):
        build_model(cfg, bad, precision="fp32")
    bad = dict(P)
    bad["head_top.weight"] = torch.zeros(7, cfg.embed_dim)
    with pytest.raises(H.HQError, match="size mismatch"):
        build_model(cfg, bad, precision="fp32")
    model = build_model(cfg, P, precision="fp32")
    with pytest.raises(ValueError, match="max_seq_len"):
        H.sampling_ihqgpt(model, 2, 0, max_seq_len=256, use_fp16=False)
    with pytest.raises(H.HQError, match="temperature"):
        H.sampling_ihqgpt(model, 2, 0, max_seq_len=64, use_fp16=False, softmax_temperature=[0.0, 1.0])
    with pytest.raises(IndexError):                       # nn.Embedding raises IndexError in the reference
        H.sampling_ihqgpt(model, 2, cfg.n_classes, max_seq_len=64, use_fp16=False)
    with pytest.raises(IndexError):
        H.sampling_ihqgpt(model, 2, 0, max_seq_len=64, use_fp16=False,
                          given_top_code=torch.full((2, 64), cfg.vocab_top, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no bf16 engine"):   # fp32-only model asked for the bf16 path
        H.sampling_ihqgpt(model, 2, 0, max_seq_len=64, use_fp16=True)


def test_shared_text_prefix_equals_per_row_prefill():
    """one prompt sampled B times - the prefill runs for one row and its KV-cache rows are broadcast
    (hq_run_args.shared_prefix); grids are identical to prefilling every row, greedy and stochastic, fp32 and bf16."""
    import hqtransformer_b200 as H
    from dataclasses import replace
    cfg = replace(O.ASYM, cond="txt")
    P = O.make_params(cfg, seed=6, init="rich")
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, cfg.vocab_txt, (1, cfg.ctx_len_txt), generator=g).repeat(7, 1)
    for precision, fp16 in (("fp32", False), ("bf16", True)):
        model = build_model(cfg, P, precision=precision, max_batch=7, max_seq_len=6)
        for kw in (dict(top_k_top=1, top_k_bot=1), dict(top_k_top=20, top_k_bot=30, softmax_temperature=[0.9, 1.1])):
            a = H.sampling_ihqgpt(model, 7, ids, max_seq_len=6, is_tqdm=False, use_fp16=fp16, seed=5, shared_prefix=False, **kw)
            b = H.sampling_ihqgpt(model, 7, ids, max_seq_len=6, is_tqdm=False, use_fp16=fp16, seed=5, shared_prefix=True, **kw)
            c = H.sampling_ihqgpt(model, 7, ids, max_seq_len=6, is_tqdm=False, use_fp16=fp16, seed=5, **kw)     # auto-detected
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])
    # greedy + one prompt: every row identical (as in the reference)
    gt, gb = H.sampling_ihqgpt(model, 7, ids, max_seq_len=6, is_tqdm=False, use_fp16=True, top_k_top=1, top_k_bot=1)
    assert bool((gt == gt[:1]).all()) and bool((gb == gb[:1]).all())
