"""Shared helpers of the parity tests: build an engine-backed model from oracle parameters."""
import json
import os

import numpy as np
import torch

from oracle import hq_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name))
    meta = json.loads(str(g["meta"]))
    return g, meta


def cfg_from_meta(meta):
    return O.HQConfig(**meta["config"])


def hparams_of(cfg: O.HQConfig, n_layers=None):
    from types import SimpleNamespace
    return SimpleNamespace(embed_dim=cfg.embed_dim, n_layers=n_layers or cfg.n_layers, n_heads=cfg.n_heads,
                           n_dense_layers=cfg.n_layers, ctx_len=None, ctx_len_img=cfg.ctx_len_img,
                           ctx_len_txt=cfg.ctx_len_txt, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, mlp_bias=True,
                           attn_bias=True, gelu_use_approx=False, use_head_txt=True, n_classes=cfg.n_classes,
                           causal_attn=None, embedding_type=getattr(cfg, "embedding_type", "transformer1"),
                           position_embedding=getattr(cfg, "position_embedding", "1d"),
                           bottom_head_type="linear", use_random_order=False, rate_random_order=1.0)


def build_model(cfg: O.HQConfig, params, precision="fp32", max_batch=8, use_cuda_graph=True, max_seq_len=64,
                use_pdl=True, use_chain=False, fuse_head_sampler=True):
    """hqtransformer_b200.iHQGPT for an oracle config, loaded with the oracle's (reference-named) parameters."""
    import hqtransformer_b200 as H
    model = H.iHQGPT(vocab_size_top=cfg.vocab_top, vocab_size_bot=cfg.vocab_bot, vocab_size_txt=cfg.vocab_txt,
                     ratio_bot2top=4, use_cls_cond=(cfg.cond == "cls"), use_txt_cond=(cfg.cond == "txt"),
                     model_type=getattr(cfg, "model_type", "parallel"), hparams=hparams_of(cfg), hparams_dec=hparams_of(cfg, cfg.n_layers_depth),
                     device=0, precision=precision, max_batch=max_batch, use_cuda_graph=use_cuda_graph,
                     max_seq_len=max_seq_len, use_pdl=use_pdl, use_chain=use_chain, fuse_head_sampler=fuse_head_sampler)
    model.load_state_dict(params, strict=True)
    return model
