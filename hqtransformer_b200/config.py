"""Stage-2 config loading, compatible with the reference's YAML schema.

The reference merges a YAML file over OmegaConf structured defaults
(hqvae/utils/config2.py:49-105, 147-163; sampling_hqmodel.py:72-75).  `omegaconf` is not a dependency
here: the same defaults are restated as plain dicts and the merge is done on PyYAML output.  Only the
fields the sampling path reads are interpreted; the rest (dataset / stage1 / optimizer / experiment) is
carried through untouched so that callers can still inspect it.
"""
from __future__ import annotations

import copy
from types import SimpleNamespace
from typing import Any, Dict, Optional

import yaml

# hqvae/utils/config2.py:49-71
STAGE2_HPARAMS_DEFAULTS: Dict[str, Any] = dict(
    embed_dim=1536, n_layers=42, n_heads=24, n_dense_layers=42, ctx_len=None, ctx_len_img=256, ctx_len_txt=64,
    embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, mlp_bias=True, attn_bias=True, gelu_use_approx=False,
    use_head_txt=True, n_classes=None, causal_attn=None, embedding_type="baseline", position_embedding="1d",
    bottom_head_type="linear", use_random_order=False, rate_random_order=1.0)

# hqvae/utils/config2.py:85-105
STAGE2_DEFAULTS: Dict[str, Any] = dict(
    type="transformer1d", vocab_size_txt=16384, vocab_size_img=16384, vocab_sizes_img=[8192, 8192, 8192],
    decoding_type=None, ratio_bot2top=4, use_pretrained=False, use_cls_cond=None, use_txt_cond=None,
    weight_bottom=4.0, weight_txt=None, weight_img=None, gamma_focal_loss=None, temp_soft_labels=None,
    use_l2norm_logits=None, hparams=None, hparams_enc=None, hparams_dec=None)


def _ns(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: _ns(v) for k, v in d.items()})
    return d


def _merge_hparams(user: Optional[dict]) -> Optional[dict]:
    if user is None:
        return None
    unknown = set(user) - set(STAGE2_HPARAMS_DEFAULTS)
    if unknown:  # OmegaConf structured merge rejects unknown keys
        raise KeyError(f"unknown stage2 hparams key(s): {sorted(unknown)}")
    out = dict(STAGE2_HPARAMS_DEFAULTS)
    out.update(user)
    return out


def merge_config(user: Dict[str, Any]) -> SimpleNamespace:
    """YAML dict -> namespace with the reference defaults filled in (get_base_config + OmegaConf.merge)."""
    user = copy.deepcopy(user or {})
    s2_user = user.get("stage2", {}) or {}
    unknown = set(s2_user) - set(STAGE2_DEFAULTS)
    if unknown:
        raise KeyError(f"unknown stage2 key(s): {sorted(unknown)}")
    s2 = dict(STAGE2_DEFAULTS)
    s2.update(s2_user)
    for k in ("hparams", "hparams_enc", "hparams_dec"):
        s2[k] = _merge_hparams(s2.get(k))
    user["stage2"] = s2
    return _ns(user)


def load_config(path: str) -> SimpleNamespace:
    with open(path, "r") as f:
        return merge_config(yaml.safe_load(f))


def engine_kwargs(config: SimpleNamespace) -> Dict[str, Any]:
    """Maps a merged config to `iHQGPT(...)` arguments the way `ImageGPT2.__init__` does
    (hqvae/models/__init__.py:123-137) and checks that the model is one this path implements."""
    s2 = config.stage2
    if "multilevel-hq" in s2.type:       # hqvae/models/__init__.py:138-145 -> HQTransformer
        return dict(vocab_sizes=list(s2.vocab_sizes_img), vocab_size_txt=s2.vocab_size_txt, decoding_type=s2.decoding_type,
                    use_cls_cond=bool(s2.use_cls_cond), use_txt_cond=bool(s2.use_txt_cond), hparams=s2.hparams,
                    hparams_dec=s2.hparams_dec)
    if "hq-transformer" not in s2.type:
        raise NotImplementedError(f"stage2.type={s2.type!r}: 'hq-transformer/*' (2-level iHQGPT) and 'multilevel-hq' "
                                  "(3-level HQTransformer) are on this path")
    model_type = s2.type.split("/")[-1] if "/" in s2.type else "top2bot"
    return dict(vocab_size_top=s2.vocab_size_img, vocab_size_bot=s2.vocab_size_img, vocab_size_txt=s2.vocab_size_txt,
                ratio_bot2top=s2.ratio_bot2top, use_cls_cond=bool(s2.use_cls_cond), use_txt_cond=bool(s2.use_txt_cond),
                model_type=model_type, hparams=s2.hparams, hparams_dec=s2.hparams_dec)
