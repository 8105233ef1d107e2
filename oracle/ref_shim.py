"""TEST INFRASTRUCTURE ONLY - imports the *unmodified* reference sampler from /root/reference.

The reference (kakaobrain/hqtransformer) is pure Python; its `hqvae.models` package pulls in
`pytorch_lightning` and `omegaconf`, neither of which is installed here.  This shim
registers bare package objects so that only the four hot-path modules are executed:

    hqvae/models/stage2/hierarchical_ar.py   (iHQGPT)
    hqvae/models/stage2/layers.py            (Block / ParallelBlock / MultiHeadSelfAttention)
    hqvae/utils/sampling.py                  (sampling_ihqgpt, cutoff_topk_logits, cutoff_topp_probs)

Nothing is copied: the modules are executed from where they lie.  `/root/reference` exists only in
the build container, so this file is used by `oracle/make_golden.py` and by the `not gpu` tests that
pin the oracle against the live reference (they skip when the tree is absent).  It must never be
imported by the product package, `bench.py` or any `-m gpu` test.
"""
import os
import sys
import types

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    """The live tree in the build container; on the GPU box the verbatim copy `oracle/install_reference.py` placed under
    the git-ignored baseline/_ref/ (used by bench.py's reference arm only)."""
    env = os.environ.get("HQ_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/hqvae"):
        return "/root/reference"
    return os.path.join(_REPO_ROOT, "baseline", "_ref")


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hqvae", "models", "stage2", "hierarchical_ar.py"))


def import_reference():
    """Returns (iHQGPT class, reference `hqvae.utils.sampling` module)."""
    if not reference_available():
        raise ImportError(f"reference tree not found under {REFERENCE_ROOT}")
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")

        class OmegaConf:  # only used in annotations by hierarchical_ar.py:15,32
            pass

        om.OmegaConf = OmegaConf
        sys.modules["omegaconf"] = om
    for name, rel in (("hqvae", "hqvae"),
                      ("hqvae.models", "hqvae/models"),
                      ("hqvae.models.stage2", "hqvae/models/stage2"),
                      ("hqvae.utils", "hqvae/utils")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, rel)]  # do NOT run the package __init__
            sys.modules[name] = pkg
    from hqvae.models.stage2.hierarchical_ar import iHQGPT  # noqa: E402
    from hqvae.utils import sampling as ref_sampling  # noqa: E402
    return iHQGPT, ref_sampling


def make_hparams(embed_dim, n_layers, n_heads, n_classes=None, ctx_len_img=64, ctx_len_txt=64,
                 embedding_type="transformer1", position_embedding="1d"):
    """Field set of hqvae/utils/config2.py:49-71 (Stage2Hparams) as a SimpleNamespace."""
    return types.SimpleNamespace(
        embed_dim=embed_dim, n_layers=n_layers, n_heads=n_heads, n_dense_layers=n_layers,
        ctx_len=None, ctx_len_img=ctx_len_img, ctx_len_txt=ctx_len_txt,
        embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, mlp_bias=True, attn_bias=True,
        gelu_use_approx=False, use_head_txt=True, n_classes=n_classes, causal_attn=None,
        embedding_type=embedding_type, position_embedding=position_embedding, bottom_head_type="linear",
        use_random_order=False, rate_random_order=1.0)


def build_reference_model(cfg, state_dict):
    """Instantiate the reference iHQGPT for an `oracle.hq_oracle.HQConfig` and load `state_dict`."""
    import contextlib
    import io
    iHQGPT, _ = import_reference()
    emb = getattr(cfg, "embedding_type", "transformer1")
    pe = getattr(cfg, "position_embedding", "1d")
    hp = make_hparams(cfg.embed_dim, cfg.n_layers, cfg.n_heads, n_classes=cfg.n_classes,
                      ctx_len_img=cfg.ctx_len_img, ctx_len_txt=cfg.ctx_len_txt, embedding_type=emb, position_embedding=pe)
    hp_dec = make_hparams(cfg.embed_dim, cfg.n_layers_depth, cfg.n_heads, n_classes=cfg.n_classes,
                          ctx_len_img=cfg.ctx_len_img, ctx_len_txt=cfg.ctx_len_txt, embedding_type=emb,
                          position_embedding=pe)
    with contextlib.redirect_stdout(io.StringIO()):
        model = iHQGPT(vocab_size_top=cfg.vocab_top, vocab_size_bot=cfg.vocab_bot,
                       vocab_size_txt=cfg.vocab_txt, ratio_bot2top=4,
                       use_cls_cond=(cfg.cond == "cls"), use_txt_cond=(cfg.cond == "txt"),
                       model_type=getattr(cfg, "model_type", "parallel"), hparams=hp, hparams_dec=hp_dec)
    missing = model.load_state_dict(state_dict, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.eval()


def reference_sample(model, num_candidates, cond, device="cpu", use_fp16=False, **kw):
    """Run the reference's own `sampling_ihqgpt` (utils/sampling.py:164-237).

    device='cpu': the reference hard-codes `.cuda()` at sampling.py:184,188; that call is neutralised for the duration
    of the run (tensor stays where it is) and fp32 is used (`use_fp16=False`).
    device='cuda': untouched - the reference in its native mode (fp16 autocast when use_fp16=True), model already on
    the GPU."""
    import torch
    _, ref_sampling = import_reference()
    if device != "cpu":
        return ref_sampling.sampling_ihqgpt(model, num_candidates=num_candidates, cond=cond, is_tqdm=False,
                                            use_fp16=use_fp16, **kw)
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        return ref_sampling.sampling_ihqgpt(model, num_candidates=num_candidates, cond=cond,
                                            is_tqdm=False, use_fp16=False, **kw)
    finally:
        torch.Tensor.cuda = saved


def build_reference_model_random(cfg):
    """The reference iHQGPT for `cfg` with its OWN random initialisation (`_init_weights`) - what `measure_throughput`
    times (config only, no checkpoint: measure_throughput/__main__.py:25-31)."""
    import contextlib
    import io
    iHQGPT, _ = import_reference()
    emb = getattr(cfg, "embedding_type", "transformer1")
    pe = getattr(cfg, "position_embedding", "1d")
    hp = make_hparams(cfg.embed_dim, cfg.n_layers, cfg.n_heads, n_classes=cfg.n_classes,
                      ctx_len_img=cfg.ctx_len_img, ctx_len_txt=cfg.ctx_len_txt, embedding_type=emb, position_embedding=pe)
    hp_dec = make_hparams(cfg.embed_dim, cfg.n_layers_depth, cfg.n_heads, n_classes=cfg.n_classes,
                          ctx_len_img=cfg.ctx_len_img, ctx_len_txt=cfg.ctx_len_txt, embedding_type=emb,
                          position_embedding=pe)
    with contextlib.redirect_stdout(io.StringIO()):
        model = iHQGPT(vocab_size_top=cfg.vocab_top, vocab_size_bot=cfg.vocab_bot, vocab_size_txt=cfg.vocab_txt,
                       ratio_bot2top=4, use_cls_cond=(cfg.cond == "cls"), use_txt_cond=(cfg.cond == "txt"),
                       model_type=getattr(cfg, "model_type", "parallel"), hparams=hp, hparams_dec=hp_dec)
    return model.eval()


# ---------------------------------------------------------------------------------------------------------------------
# stage 1: the unmodified `SimRQGAN2Generator` for the decode_code oracle pin
# ---------------------------------------------------------------------------------------------------------------------
class _AttrDict(dict):
    """Stands in for the OmegaConf node: `Encoder(**hparams)` needs a mapping, the generator reads attributes."""
    __getattr__ = dict.__getitem__


def build_reference_stage1(cfg, state_dict):
    """Reference `SimRQGAN2Generator` (hqvae/models/stage1/generator.py:176-260) for an `oracle.s1_oracle.S1Config`, with the
    decode-side parameters of `state_dict` loaded (the encoder keeps its own random init: decode_code never touches it)."""
    import_reference()
    for name, rel in (("hqvae.models.stage1", "hqvae/models/stage1"),
                      ("hqvae.models.stage1.modules", "hqvae/models/stage1/modules")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
            sys.modules[name] = pkg
    from hqvae.models.stage1.generator import SimRQGAN2Generator  # noqa: E402
    hp = _AttrDict(double_z=False, z_channels=cfg.z_channels, resolution=cfg.resolution, in_channels=3, out_ch=cfg.out_ch,
                   ch=cfg.ch, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                   attn_resolutions=list(cfg.attn_resolutions), pdrop=0.0, use_init_downsample=True, use_mid_block=True,
                   use_attn=True)
    aux = types.SimpleNamespace(shared_codebook=False, decoding_type="concat", upsample="pixelshuffle", bottom_start=0,
                                bottom_window=2)
    model = SimRQGAN2Generator(n_embed=cfg.n_embed, embed_dim=cfg.embed_dim, ema_update=True, hparams=hp, hparams_aux=aux)
    res = model.load_state_dict(state_dict, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith(("encoder.", "quant_conv_b.", "quantize_t.cluster", "quantize_t.embedding_avg",
                             "quantize_b.cluster", "quantize_b.embedding_avg")) for k in res.missing_keys), res.missing_keys
    return model.eval()


# ---------------------------------------------------------------------------------------------------------------------
# 3-level HQTransformer
# ---------------------------------------------------------------------------------------------------------------------
def build_reference_hq3(cfg, state_dict):
    """Reference `HQTransformer` (hqvae/models/stage2/hqtransformer.py:168-216), decoding_type 'parallel-add', for an
    `oracle.hq3_oracle.HQ3Config`."""
    import contextlib
    import io
    import_reference()
    from hqvae.models.stage2.hqtransformer import HQTransformer  # noqa: E402
    hp = make_hparams(cfg.embed_dim, cfg.n_layers, cfg.n_heads, n_classes=cfg.n_classes, ctx_len_img=cfg.ctx_len_img)
    hp_dec = make_hparams(cfg.embed_dim, cfg.n_layers_depth, cfg.n_heads, n_classes=cfg.n_classes, ctx_len_img=cfg.ctx_len_img)
    with contextlib.redirect_stdout(io.StringIO()):
        model = HQTransformer(vocab_sizes=list(cfg.vocab_sizes), vocab_size_txt=16, decoding_type="parallel-add",
                              use_cls_cond=(cfg.cond == "cls"), use_txt_cond=False, hparams=hp, hparams_dec=hp_dec)
    res = model.load_state_dict(state_dict, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model.eval()


def reference_sample_hq3(model, num_candidates, cond, **kw):
    """The reference's own `sampling_hqtransformer` (utils/sampling.py:240-307) on CPU (its `.cuda()` calls neutralised)."""
    import torch
    _, ref_sampling = import_reference()
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        return ref_sampling.sampling_hqtransformer(model, num_candidates=num_candidates, cond=cond, is_tqdm=False,
                                                   use_fp16=False, **kw)
    finally:
        torch.Tensor.cuda = saved
