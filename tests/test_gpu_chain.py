"""-m gpu: the persistent chain kernel (csrc/chain.cuh: one launch per run of dependent GEMM / LayerNorm / depth-attention
ops, grid barrier between ops) against the one-kernel-per-op path.  Both paths accumulate every output element in the
same order (same K order, same split-K slices, same LayerNorm reductions), so the comparison is EXACT."""
import pytest
import torch

from oracle import hq_oracle as O
from tests.helpers import build_model

pytestmark = pytest.mark.gpu


def _labels(cfg, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, cfg.n_classes, (B,), generator=g)


@pytest.mark.parametrize("cfg_name,B", [("SMALL", 150), ("ASYM", 129), ("TINY", 300), ("SMALL", 600)])
def test_chain_logits_equal_per_op_kernels(cfg_name, B):
    """Teacher-forced head outputs of every position, chain kernel vs. per-op kernels: bit-identical (stream launches)."""
    import hqtransformer_b200 as H
    cfg = getattr(O, cfg_name)
    P = O.make_params(cfg, seed=11, init="rich")
    labels = _labels(cfg, B)
    S = 6
    g = torch.Generator().manual_seed(1)
    ct = torch.randint(0, cfg.vocab_top, (B, S), generator=g)
    cb = torch.randint(0, cfg.vocab_bot, (B, S, 4), generator=g)
    ref = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S, use_chain=False)
    want = H.step_logits(ref, labels, ct, cb, use_fp16=True)
    assert ref.engine("bf16").chain_launches == 0
    del ref
    model = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S, use_chain=True)
    got = H.step_logits(model, labels, ct, cb, use_fp16=True)
    assert model.engine("bf16").chain_launches > 0, "the chain kernel did not run"
    assert torch.isfinite(got).all()
    assert torch.equal(got, want), float((got - want).abs().max())


@pytest.mark.parametrize("graph,pdl", [(False, False), (True, False), (False, True), (True, True)])
def test_chain_sampling_equal_per_op_kernels_all_launch_modes(graph, pdl):
    """Whole sampling runs (greedy and stochastic, 16 positions, B = 300), repeated: the chain path under every launch
    mode reproduces the per-op path's grids exactly (a grid-barrier or PDL race would show up as a mismatch)."""
    import hqtransformer_b200 as H
    cfg = O.SMALL
    P = O.make_params(cfg, seed=5, init="rich")
    B = 300
    labels = _labels(cfg, B).cuda()
    ref = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=16, use_cuda_graph=False, use_pdl=False,
                      use_chain=False)
    model = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=16, use_cuda_graph=graph, use_pdl=pdl,
                        use_chain=True)
    for kw in (dict(top_k_top=1, top_k_bot=1),
               dict(top_k_top=50, top_p_top=0.9, top_k_bot=50, top_p_bot=0.9, softmax_temperature=[0.9, 0.9])):
        ct0, cb0 = H.sampling_ihqgpt(ref, B, labels, max_seq_len=16, is_tqdm=False, use_fp16=True, seed=3, **kw)
        for _ in range(3):
            ct, cb = H.sampling_ihqgpt(model, B, labels, max_seq_len=16, is_tqdm=False, use_fp16=True, seed=3, **kw)
            assert torch.equal(ct, ct0) and torch.equal(cb, cb0), kw
    assert model.engine("bf16").chain_launches > 0


def test_chain_text_prefix_model():
    """Text-conditional model: the 64-token causal prefill stays on the per-op path, every decode position takes the
    chain kernel (cache slots T0 + pos - 1 come in as a per-launch value)."""
    import hqtransformer_b200 as H
    from dataclasses import replace
    cfg = replace(O.ASYM, cond="txt")
    P = O.make_params(cfg, seed=3, init="rich")
    B = 140
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(0, cfg.vocab_txt, (B, cfg.ctx_len_txt), generator=g).cuda()
    kw = dict(top_k_top=20, top_k_bot=20, softmax_temperature=[0.9, 0.9], max_seq_len=8, is_tqdm=False, use_fp16=True, seed=9)
    ref = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=8, use_chain=False)
    ct0, cb0 = H.sampling_ihqgpt(ref, B, ids, **kw)
    del ref
    model = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=8, use_chain=True)
    for _ in range(2):
        ct, cb = H.sampling_ihqgpt(model, B, ids, **kw)
        assert torch.equal(ct, ct0) and torch.equal(cb, cb0)
    assert model.engine("bf16").chain_launches > 0


def test_chain_per_position_api_and_small_batches():
    """sampling_step-style runs over [pos, pos + 1) use the chain too; batches <= 128 rows stay on the per-op path."""
    import hqtransformer_b200 as H
    from hqtransformer_b200.engine import SamplingParams
    cfg = O.SMALL
    P = O.make_params(cfg, seed=5, init="rich")
    B, S = 160, 5
    labels = _labels(cfg, B).cuda()
    model = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S, use_chain=True)
    eng = model.engine("bf16")
    sp = SamplingParams(top_k_top=30, top_k_bot=30, seed=4)
    ct = torch.zeros(B, S, dtype=torch.int64, device="cuda")
    cb = torch.zeros(B, S, 4, dtype=torch.int64, device="cuda")
    eng.run(batch=B, seq_len=S, pos_begin=0, pos_end=S, sampling=sp, cond=labels, codes_top=ct, codes_bot=cb)
    ct2, cb2 = torch.zeros_like(ct), torch.zeros_like(cb)
    for p in range(S):
        eng.run(batch=B, seq_len=S, pos_begin=p, pos_end=p + 1, sampling=sp, cond=labels, codes_top=ct2, codes_bot=cb2)
    assert torch.equal(ct, ct2) and torch.equal(cb, cb2)
    before = eng.chain_launches
    eng.run(batch=64, seq_len=S, pos_begin=0, pos_end=S, sampling=sp, cond=labels[:64], codes_top=ct[:64].contiguous(),
            codes_bot=cb[:64].contiguous())
    assert eng.chain_launches == before
