"""-m gpu: the 3-level HQTransformer sampler (hq_config.code_levels = 3) through the C ABI against the
reference-made golden and the oracle."""
import numpy as np
import pytest
import torch

from oracle import hq3_oracle as O3
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _model(cfg, P, precision, max_batch, max_seq_len=64, **kw):
    import hqtransformer_b200 as H
    from types import SimpleNamespace

    def hp(n_layers):
        return SimpleNamespace(embed_dim=cfg.embed_dim, n_layers=n_layers, n_heads=cfg.n_heads, n_dense_layers=n_layers,
                               ctx_len=None, ctx_len_img=cfg.ctx_len_img, ctx_len_txt=64, n_classes=cfg.n_classes,
                               embedding_type="transformer1", position_embedding="1d", gelu_use_approx=False)
    m = H.HQTransformer(vocab_sizes=list(cfg.vocab_sizes), vocab_size_txt=16, decoding_type="parallel-add",
                        use_cls_cond=(cfg.cond == "cls"), use_txt_cond=False, hparams=hp(cfg.n_layers),
                        hparams_dec=hp(cfg.n_layers_depth), precision=precision, max_batch=max_batch, max_seq_len=max_seq_len, **kw)
    assert list(m.param_shapes().keys()) == list(O3.param_shapes(cfg).keys())
    m.load_state_dict(P, strict=True)
    return m


@pytest.mark.parametrize("graph,pdl", [(False, False), (True, True)])
def test_level3_greedy_codes_bit_exact_vs_reference_fp32(graph, pdl):
    """Greedy grids of the unmodified reference (CPU fp32; 1512 decisions, margin >= 2e-4) reproduced bit for bit by the fp32 engine."""
    import hqtransformer_b200 as H
    g, meta = load_golden("tiny3_cls_greedy.npz")
    cfg = O3.HQ3Config.from_dict(meta["config"])
    P = O3.make_params(cfg, seed=meta["seed"])
    labels = torch.from_numpy(g["labels"])
    S = g["codes_top"].shape[1]
    m = _model(cfg, P, "fp32", len(labels), use_cuda_graph=graph, use_pdl=pdl)
    ct, cm, cb = H.sampling_hqtransformer(m, len(labels), labels, top_k=[1, 1, 1], top_p=[1.0, 1.0, 1.0], use_fp16=False,
                                          max_seq_len=S, is_tqdm=False)
    assert tuple(cm.shape) == (len(labels), S, 4) and tuple(cb.shape) == (len(labels), S, 16)
    assert np.array_equal(ct.cpu().numpy(), g["codes_top"])
    assert np.array_equal(cm.cpu().numpy(), g["codes_mid"])
    assert np.array_equal(cb.cpu().numpy(), g["codes_bot"])


@pytest.mark.parametrize("cfg_name,B", [("TINY3", 5), ("SMALL3", 4), ("SMALL3", 20)])
def test_level3_step_logits_vs_oracle(cfg_name, B):
    """Teacher-forced head outputs of all 21 stack slots: fp32 engine vs fp32 oracle (< 2e-5); bf16 engine vs the
    rounding-emulating oracle (max-abs <= 2e-2, mean <= 2e-3).  B = 20 -> 320 rows in the bottom pass (CTA-pair GEMMs)."""
    import hqtransformer_b200 as H
    cfg = getattr(O3, cfg_name)
    P = O3.make_params(cfg, seed=7)
    S = 3
    g = torch.Generator().manual_seed(B)
    labels = torch.randint(0, cfg.n_classes, (B,), generator=g)
    given = torch.stack([torch.randint(0, cfg.vocab_sizes[0 if j == 0 else (1 if j < 5 else 2)], (B, S), generator=g)
                         for j in range(21)], dim=-1)
    codes = [given[:, :, 0], given[:, :, 1:5].contiguous(), given[:, :, 5:].contiguous()]
    if B <= 8:
        _, want = O3.sample(P, cfg, labels, B, max_seq_len=S, given=given, return_logits=True)
        m32 = _model(cfg, P, "fp32", B, max_seq_len=S)
        got = H.step_logits3(m32, labels, codes, use_fp16=False).cpu()
        assert float((got - want).abs().max()) < 2e-5
    _, emu = O3.sample(P, cfg, labels, B, max_seq_len=S, given=given, return_logits=True, emulate="bf16")
    m16 = _model(cfg, P, "bf16", B, max_seq_len=S)
    lg = H.step_logits3(m16, labels, codes, use_fp16=True).cpu()
    err = (lg - emu).abs()
    assert float(err.max()) <= 2e-2 and float(err.mean()) <= 2e-3, (float(err.max()), float(err.mean()))


def test_level3_stochastic_sampling_deterministic_and_sharding_invariant():
    import hqtransformer_b200 as H
    cfg = O3.SMALL3
    P = O3.make_params(cfg, seed=5)
    B = 40
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, cfg.n_classes, (B,), generator=g).cuda()
    m = _model(cfg, P, "bf16", B, max_seq_len=6)
    kw = dict(top_k=[50, 40, 30], top_p=[0.9, 0.95, 0.9], softmax_temperature=[0.9, 1.0, 1.1], max_seq_len=6, is_tqdm=False,
              use_fp16=True, seed=3)
    a = H.sampling_hqtransformer(m, B, labels, **kw)
    b = H.sampling_hqtransformer(m, B, labels, **kw)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    for lvl, x in enumerate(a):
        assert int(x.min()) >= 0 and int(x.max()) < cfg.vocab_sizes[lvl]
    s = H.sampling_hqtransformer(m, 10, labels[30:40], row_offset=30, **kw)
    assert all(torch.equal(x, y[30:40]) for x, y in zip(s, a))


def test_level3_full_size_architecture_runs():
    """ImageNet level-3 architecture (D = 1536, 12 + 4 layers, 3 x 8192 codes), random init, a short run at B = 32."""
    import hqtransformer_b200 as H
    cfg = O3.IMAGENET_L12_LEVEL3
    from types import SimpleNamespace
    hp = SimpleNamespace(embed_dim=cfg.embed_dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, ctx_len_img=256, n_classes=1000,
                         embedding_type="transformer1", position_embedding="1d")
    m = H.HQTransformer(vocab_sizes=[8192, 8192, 8192], vocab_size_txt=16384, decoding_type="parallel-add", use_cls_cond=True,
                        use_txt_cond=False, hparams=hp, hparams_dec=None, precision="bf16", max_batch=32, max_seq_len=4)
    m.init_weights(seed=0)
    ct, cm, cb = H.sampling_hqtransformer(m, 32, 7, max_seq_len=4, is_tqdm=False, seed=1)
    assert tuple(ct.shape) == (32, 4) and tuple(cm.shape) == (32, 4, 4) and tuple(cb.shape) == (32, 4, 16)
    assert int(cb.max()) < 8192 and int(cb.min()) >= 0
