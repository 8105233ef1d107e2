"""-m gpu: the head GEMM that draws in its epilogue (csrc/gemm.cuh EPI_SAMPLE + sample_finalize_kernel; VERDICT n5).

Draws with top_k None / top_p None (the measure_throughput protocol) never write the [rows, V] logits: each 32-column chunk
of a row yields (log-sum-exp, one index drawn inside the chunk), a second kernel draws the chunk.  Checked here:
  * the empirical distribution equals softmax(logits / T) of the SAME engine's teacher-forced logits (chi-square), for the
    top draw and for each of the four bottom draws, with temperatures != 1;
  * results do not depend on the kernel shape (CTA pairs at M > 128, single CTAs below), the batch split (row_offset),
    CUDA graphs or PDL - the property multi-GPU sharding relies on;
  * draws with a top-k / top-p cut, greedy draws and the fp32 engine still take the unfused sampler.
"""
import numpy as np
import pytest
import torch

from oracle import hq_oracle as O
from tests.helpers import build_model

pytestmark = pytest.mark.gpu


def _chi2_ok(counts: np.ndarray, pr: np.ndarray, n: int):
    """Pearson chi-square with every class of expectation < 5 pooled into one bucket; 5-sigma bound on the statistic."""
    exp = pr * n
    keep = exp >= 5
    obs = np.concatenate([counts[keep], [counts[~keep].sum()]])
    ex = np.concatenate([exp[keep], [exp[~keep].sum()]])
    ok = ex > 0
    chi2 = (((obs - ex) ** 2)[ok] / ex[ok]).sum()
    dof = max(int(ok.sum()) - 1, 1)
    return chi2 < dof + 5 * (2 * dof) ** 0.5, (float(chi2), dof)


def _softmax(z: np.ndarray, t: float) -> np.ndarray:
    z = z.astype(np.float64) / t
    z -= z.max()
    p = np.exp(z)
    return p / p.sum()


def test_fused_draws_follow_the_engines_own_logits():
    import hqtransformer_b200 as H
    cfg = O.SMALL
    P = O.make_params(cfg, seed=21, init="rich")
    B = 16384
    temps = [0.8, 1.25]
    model = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=1)
    labels = torch.full((B,), 3, dtype=torch.int64)
    ct, cb = H.sampling_ihqgpt(model, B, labels, use_fp16=True, max_seq_len=1, is_tqdm=False, seed=5,
                               softmax_temperature=temps)
    ct, cb = ct.cpu().numpy(), cb.cpu().numpy()
    # top draw: every row sees the same logits
    lg = H.step_logits(model, labels[:4], torch.zeros(4, 1, dtype=torch.int64), torch.zeros(4, 1, 4, dtype=torch.int64), use_fp16=True)
    pr = _softmax(lg[0, 0, 0, :cfg.vocab_top].cpu().numpy(), temps[0])
    ok, info = _chi2_ok(np.bincount(ct[:, 0], minlength=cfg.vocab_top).astype(np.float64), pr, B)
    assert ok, ("top", info)
    # bottom draws: force one top code for every row, then the four bottom rows of all images see the same logits
    c_star = int(np.bincount(ct[:, 0]).argmax())
    given = torch.full((B, 1), c_star, dtype=torch.int64)
    ct2, cb2 = H.sampling_ihqgpt(model, B, labels, use_fp16=True, max_seq_len=1, is_tqdm=False, seed=6,
                                 softmax_temperature=temps, given_top_code=given)
    assert (ct2.cpu().numpy() == c_star).all()
    cb2 = cb2.cpu().numpy()
    lg2 = H.step_logits(model, labels[:4], given[:4], torch.zeros(4, 1, 4, dtype=torch.int64), use_fp16=True)
    for j in range(4):
        prj = _softmax(lg2[0, 0, 1 + j, :cfg.vocab_bot].cpu().numpy(), temps[1])
        ok, info = _chi2_ok(np.bincount(cb2[:, 0, j], minlength=cfg.vocab_bot).astype(np.float64), prj, B)
        assert ok, ("bottom", j, info)
    # the four bottom slots are distinct draws
    assert not (cb2[:, 0, 0] == cb2[:, 0, 1]).all()
    # the unfused sampler draws from the same distribution with another stream
    ref = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=1, fuse_head_sampler=False)
    ct3, _ = H.sampling_ihqgpt(ref, B, labels, use_fp16=True, max_seq_len=1, is_tqdm=False, seed=5, softmax_temperature=temps)
    ct3 = ct3.cpu().numpy()
    assert not np.array_equal(ct3, ct)
    ok, info = _chi2_ok(np.bincount(ct3[:, 0], minlength=cfg.vocab_top).astype(np.float64), pr, B)
    assert ok, ("unfused top", info)


@pytest.mark.parametrize("cfg_name", ["SMALL", "ASYM"])
def test_fused_draws_do_not_depend_on_kernel_shape_sharding_or_launch_mode(cfg_name):
    import hqtransformer_b200 as H
    cfg = getattr(O, cfg_name)
    P = O.make_params(cfg, seed=9, init="rich")
    B, S = 600, 4
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, cfg.n_classes, (B,), generator=g)
    kw = dict(use_fp16=True, max_seq_len=S, is_tqdm=False, seed=77, softmax_temperature=[0.9, 1.1])
    model = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S)
    ct, cb = H.sampling_ihqgpt(model, B, labels, **kw)
    ct1, cb1 = H.sampling_ihqgpt(model, B, labels, **kw)
    assert torch.equal(ct, ct1) and torch.equal(cb, cb1)                      # deterministic
    ctd, _ = H.sampling_ihqgpt(model, B, labels, **dict(kw, seed=78))
    assert not torch.equal(ctd, ct)
    # shards: 100 rows (single-CTA GEMM kernel, M <= 128), 300 rows (CTA pairs, other tile counts)
    for lo, n in ((0, 100), (100, 100), (200, 300), (500, 100)):
        cts, cbs = H.sampling_ihqgpt(model, n, labels[lo:lo + n], row_offset=lo, **kw)
        assert torch.equal(cts, ct[lo:lo + n]) and torch.equal(cbs, cb[lo:lo + n]), (lo, n)
    del model
    for graph, pdl in ((False, False), (True, False), (False, True)):
        m2 = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S, use_cuda_graph=graph, use_pdl=pdl)
        for _ in range(2):
            ct2, cb2 = H.sampling_ihqgpt(m2, B, labels, **kw)
            assert torch.equal(ct2, ct) and torch.equal(cb2, cb), (graph, pdl)
        del m2


def test_filtered_greedy_and_fp32_runs_keep_the_unfused_sampler():
    """A cut on one level only fuses the other level's head; greedy and fp32 runs are bit-identical with fusion on / off."""
    import hqtransformer_b200 as H
    cfg = O.SMALL
    P = O.make_params(cfg, seed=9, init="rich")
    B, S = 300, 3
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, cfg.n_classes, (B,), generator=g)
    on = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S)
    off = build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=S, fuse_head_sampler=False)
    base = dict(use_fp16=True, max_seq_len=S, is_tqdm=False, seed=3)
    for kw in (dict(top_k_top=1, top_k_bot=1), dict(top_k_top=40, top_k_bot=40, top_p_top=0.9, top_p_bot=0.9),
               dict(top_k_top=40, top_p_bot=0.5)):
        a = H.sampling_ihqgpt(on, B, labels, **base, **kw)
        b = H.sampling_ihqgpt(off, B, labels, **base, **kw)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), kw
    # top filtered, bottom free: the top codes (unfused on both) agree at position 0, the bottom draws come from the fused head
    a = H.sampling_ihqgpt(on, B, labels, **base, top_k_top=40)
    b = H.sampling_ihqgpt(off, B, labels, **base, top_k_top=40)
    assert torch.equal(a[0][:, 0], b[0][:, 0]) and not torch.equal(a[1], b[1])
    del on, off
    f_on = build_model(cfg, P, precision="fp32", max_batch=32, max_seq_len=S)
    f_off = build_model(cfg, P, precision="fp32", max_batch=32, max_seq_len=S, fuse_head_sampler=False)
    a = H.sampling_ihqgpt(f_on, 32, labels[:32], use_fp16=False, max_seq_len=S, is_tqdm=False, seed=3)
    b = H.sampling_ihqgpt(f_off, 32, labels[:32], use_fp16=False, max_seq_len=S, is_tqdm=False, seed=3)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_fused_draws_three_level_model():
    """The 3-level HQTransformer: all three heads (1 + 4 + 16 rows per image) draw in the epilogue; shard-invariant."""
    import hqtransformer_b200 as H
    from oracle import hq3_oracle as O3
    from tests.test_gpu_level3 import _model as build3
    cfg = O3.TINY3
    P = O3.make_params(cfg, seed=4)
    B, S = 200, 3
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, cfg.n_classes, (B,), generator=g)
    model = build3(cfg, P, "bf16", B, max_seq_len=S)
    kw = dict(use_fp16=True, max_seq_len=S, is_tqdm=False, seed=12, softmax_temperature=[0.9, 1.0, 1.1])
    codes = H.sampling_hqtransformer(model, B, labels, **kw)
    again = H.sampling_hqtransformer(model, B, labels, **kw)
    assert all(torch.equal(a, b) for a, b in zip(codes, again))
    shard = H.sampling_hqtransformer(model, 60, labels[100:160], row_offset=100, **kw)
    assert all(torch.equal(a, b[100:160]) for a, b in zip(shard, codes))
    for c, v in zip(codes, cfg.vocab_sizes):
        assert int(c.min()) >= 0 and int(c.max()) < v
