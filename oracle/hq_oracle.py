"""TEST INFRASTRUCTURE ONLY - CPU oracle for the HQ-Transformer 2-level sampling loop.

This is a from-scratch restatement (torch CPU tensors, fp32, functional style, no nn.Module) of the
algorithm in the reference's hot path.  It exists to CHECK the CUDA path; it is never the product
and never a fallback.  Only `tests/`, `__graft_entry__.smoke()`, and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.

Pinning status: the reference ships no tests or golden vectors, so parity is
pinned by running the *unmodified reference itself* (oracle/ref_shim.py) in the build container:
`oracle/make_golden.py` stores its outputs under tests/golden/, and tests/test_oracle_vs_reference.py
re-checks this file against the live reference whenever /root/reference is present.

Each function cites the reference lines (relative to /root/reference/) it restates.

`emulate="bf16"` reproduces the rounding points of the CUDA production path (bf16 weights, bf16
GEMM-input activations, bf16 q/k/v and KV cache, fp32 accumulation/residual/LayerNorm/softmax) so
that the tcgen05 kernels can be compared far more tightly than a plain bf16-vs-fp32 tolerance.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
from torch.nn import functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# configuration / parameters
# ----------------------------------------------------------------------------------------------
@dataclass
class HQConfig:
    """Architecture of an `iHQGPT(model_type='parallel', embedding_type='transformer1')`
    (hqvae/models/stage2/hierarchical_ar.py:24-216; YAML fields hqvae/utils/config2.py:49-105)."""
    embed_dim: int = 1536
    n_heads: int = 24
    n_layers: int = 12
    n_layers_depth: int = 4          # hparams_dec absent -> 4 (hierarchical_ar.py:150-153)
    vocab_top: int = 8192
    vocab_bot: int = 8192
    vocab_txt: int = 16384
    n_classes: int = 1000
    ctx_len_img: int = 256           # rows of pos_emb_top; only 0..63 are used when sampling 8x8
    ctx_len_txt: int = 64
    cond: str = "cls"                # 'cls' | 'txt' | 'uncond'
    # variants of the 2-level model; the defaults are the ImageNet / CC-15M checkpoints' settings
    embedding_type: str = "transformer1"   # 'transformer1' | 'reduce' (FFHQ checkpoint; hierarchical_ar.py:85-88, 522-526)
    position_embedding: str = "1d"         # '1d' | '2d' (hierarchical_ar.py:118-125, 506-514)
    model_type: str = "parallel"           # 'parallel' | 'top2bot' (:565-664) | 'bidirectional' (:791-878)

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.n_heads

    @property
    def idx_pred(self) -> int:       # hierarchical_ar.py:66,74,78
        return self.ctx_len_txt if self.cond == "txt" else 0

    def to_dict(self):
        return asdict(self)


IMAGENET_L12 = HQConfig()
IMAGENET_L24 = HQConfig(n_layers=24)
IMAGENET_L42 = HQConfig(n_layers=42, n_layers_depth=6)
CC15M_L12 = HQConfig(cond="txt", ctx_len_img=64, n_classes=1000)
SMALL = HQConfig(embed_dim=256, n_heads=4, n_layers=4, n_layers_depth=4, vocab_top=1024, vocab_bot=1024,
                 vocab_txt=512, n_classes=10, ctx_len_img=64, ctx_len_txt=64)
TINY = HQConfig(embed_dim=128, n_heads=2, n_layers=2, n_layers_depth=2, vocab_top=256, vocab_bot=256,
                vocab_txt=128, n_classes=10, ctx_len_img=64, ctx_len_txt=64)
# nothing symmetric: spatial depth != depth-transformer depth (hparams_dec given explicitly, as in the L42 yaml),
# top vocabulary != bottom vocabulary, 6 heads (2 per attention work item)
ASYM = HQConfig(embed_dim=384, n_heads=6, n_layers=3, n_layers_depth=2, vocab_top=512, vocab_bot=768,
                vocab_txt=128, n_classes=7, ctx_len_img=64, ctx_len_txt=64)


def _block_shapes(prefix: str, D: int) -> "OrderedDict[str, Tuple[int, ...]]":
    """state_dict entries of one Block / ParallelBlock (layers.py:290-317, 332-364) in module order."""
    s = OrderedDict()
    s[f"{prefix}.ln1.weight"] = (D,)
    s[f"{prefix}.ln1.bias"] = (D,)
    s[f"{prefix}.ln2.weight"] = (D,)
    s[f"{prefix}.ln2.bias"] = (D,)
    for nm in ("key", "query", "value", "proj"):          # layers.py:43-52
        s[f"{prefix}.attn.{nm}.weight"] = (D, D)
        s[f"{prefix}.attn.{nm}.bias"] = (D,)
    s[f"{prefix}.mlp.0.weight"] = (4 * D, D)              # layers.py:312-317
    s[f"{prefix}.mlp.0.bias"] = (4 * D,)
    s[f"{prefix}.mlp.2.weight"] = (D, 4 * D)
    s[f"{prefix}.mlp.2.bias"] = (D,)
    return s


def param_shapes(cfg: HQConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Every state_dict key of the reference iHQGPT for this config, with its shape
    (hierarchical_ar.py:63-209).  `load_state_dict(strict=True)` on the reference accepts exactly this set."""
    D = cfg.embed_dim
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    if cfg.cond == "cls":
        s["sos.weight"] = (cfg.n_classes, D)              # :65
    elif cfg.cond == "txt":
        s["tok_emb_txt.weight"] = (cfg.vocab_txt, D)      # :68-72
        s["pos_emb_txt.weight"] = (cfg.ctx_len_txt, D)
        s["head_txt.weight"] = (cfg.vocab_txt, D)
        s["ln_txt.weight"] = (D,)
        s["ln_txt.bias"] = (D,)
    else:
        s["sos"] = (1, 1, D)                              # :77
    s["sos_depth"] = (1, 1, D)                            # :156
    s["tok_emb_top.weight"] = (cfg.vocab_top, D)          # :101-103
    if cfg.embedding_type == "reduce":                    # :85-88: bottom embeddings are D/4 wide, no pos_emb_emb
        s["tok_emb_bot.weight"] = (cfg.vocab_bot, D // 4)
    else:
        s["tok_emb_bot.weight"] = (cfg.vocab_bot, D)
        s["pos_emb_emb.weight"] = (5, D)
    if cfg.position_embedding == "2d":                    # :121-125
        H = int(math.sqrt(cfg.ctx_len_img))
        s["pos_emb_top_h.weight"] = (H, D)
        s["pos_emb_top_w.weight"] = (H, D)
    else:
        s["pos_emb_top.weight"] = (cfg.ctx_len_img, D)    # :120
    for i in range(cfg.n_layers):                         # :134-142
        s.update(_block_shapes(f"blocks.{i}", D))
    s["ln_f.weight"] = (D,)                               # :144
    s["ln_f.bias"] = (D,)
    s["tok_emb_top_depth.weight"] = (cfg.vocab_top, D)    # :160
    s["tok_emb_bot_depth.weight"] = (cfg.vocab_bot, D)    # :165 (never read when sampling 'parallel')
    s["pos_emb_depth.weight"] = (5, D)                    # :167
    for i in range(cfg.n_layers_depth):                   # :174-182
        s.update(_block_shapes(f"depths.{i}", D))
    s["ln_top.weight"] = (D,)                             # :205-209
    s["ln_top.bias"] = (D,)
    s["head_top.weight"] = (cfg.vocab_top, D)
    s["ln_bot.weight"] = (D,)
    s["ln_bot.bias"] = (D,)
    s["head_bot.weight"] = (cfg.vocab_bot, D)
    return s


def make_params(cfg: HQConfig, seed: int = 0, init: str = "reference",
                device: str = "cpu") -> "OrderedDict[str, Tensor]":
    """Random-init weights, deterministic in (cfg, seed, init) for one torch build.

    init='reference' follows `iHQGPT._init_weights` (hierarchical_ar.py:218-225): Linear/Embedding
    weights N(0, 0.02), biases 0, LayerNorm (1, 0); the raw nn.Parameters `sos_depth` / uncond `sos`
    keep their `torch.randn` values (:77, :156).
    init='rich' additionally randomises biases and LayerNorm affines (as a trained checkpoint would
    have) so that every bias/affine code path influences the result; used by the parity tests."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        leaf = name.split(".")[-1]
        is_ln = (".ln" in name or name.startswith("ln_")) and "head" not in name
        if name in ("sos", "sos_depth"):
            t = torch.randn(shape, generator=g)
        elif is_ln:
            if init == "rich":
                t = (1.0 + 0.1 * torch.randn(shape, generator=g)) if leaf == "weight" \
                    else 0.05 * torch.randn(shape, generator=g)
            else:
                t = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        elif leaf == "bias":
            t = 0.02 * torch.randn(shape, generator=g) if init == "rich" else torch.zeros(shape)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.to(device)
    return out


# ----------------------------------------------------------------------------------------------
# rounding emulation
# ----------------------------------------------------------------------------------------------
def _identity(t: Tensor) -> Tensor:
    return t


def _round_bf16(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


_GEMM_WEIGHT_LEAVES = ("attn.key.weight", "attn.query.weight", "attn.value.weight", "attn.proj.weight",
                       "mlp.0.weight", "mlp.2.weight")


def round_params(params: Dict[str, Tensor], emulate: Optional[str]) -> Dict[str, Tensor]:
    """bf16 mode of the CUDA path stores the GEMM weights (q/k/v/proj/mlp, both heads) in bf16;
    embedding tables, biases, LayerNorm affines and sos_depth stay fp32 (as under the reference's
    autocast, where only nn.Linear/bmm run in half precision)."""
    if emulate is None:
        return params
    assert emulate == "bf16"
    out = OrderedDict()
    for k, v in params.items():
        if k.endswith(_GEMM_WEIGHT_LEAVES) or k in ("head_top.weight", "head_bot.weight"):
            out[k] = _round_bf16(v)
        else:
            out[k] = v
    return out


# ----------------------------------------------------------------------------------------------
# filters  (hqvae/utils/sampling.py)
# ----------------------------------------------------------------------------------------------
def cutoff_topk_logits(logits: Tensor, k: Optional[int]) -> Tensor:
    """sampling.py:12-19 - keep every logit >= the k-th largest (ties kept), the rest -> -inf."""
    if k is None:
        return logits
    kth = torch.topk(logits, k, dim=-1).values[:, -1:]
    return torch.where(logits < kth, torch.full_like(logits, -float("inf")), logits)


def cutoff_topp_probs(probs: Tensor, p: Optional[float]) -> Tensor:
    """sampling.py:22-37 - nucleus: in descending order, entry j>0 is dropped iff the cumulative mass of
    entries 0..j-1 is already >= p (entry 0 is always kept); survivors are renormalised."""
    if p is None:
        return probs
    sorted_probs, order = torch.sort(probs, dim=-1, descending=True)
    cum = torch.cumsum(sorted_probs, dim=-1)
    drop_sorted = torch.zeros_like(cum, dtype=torch.bool)
    drop_sorted[..., 1:] = cum[..., :-1] >= p
    drop = torch.zeros_like(drop_sorted).scatter(-1, order, drop_sorted)
    kept = probs.masked_fill(drop, 0.0)
    return kept / kept.sum(dim=-1, keepdim=True)


def get_positional_encoding(inputs: Tensor, mode: str = "1d") -> Tensor:
    """sampling.py:40-52 ('1d' only: arange(N) repeated over the batch)."""
    if mode != "1d":
        raise ValueError("%s positional encoding invalid" % mode)
    B, N = inputs.shape
    return torch.arange(N, device=inputs.device).repeat((B, 1))


def draw_token(logits: Tensor, temperature: float, top_k: Optional[int], top_p: Optional[float],
               generator: Optional[torch.Generator] = None) -> Tuple[Tensor, Tensor]:
    """hierarchical_ar.py:763-769 / 779-784: z/=T; top-k; softmax; top-p; multinomial(1).
    Returns (idx [R,1] int64, probs [R,V])."""
    z = logits / temperature
    z = cutoff_topk_logits(z, top_k)
    probs = F.softmax(z, dim=-1)
    probs = cutoff_topp_probs(probs, top_p)
    idx = torch.multinomial(probs, num_samples=1, generator=generator)
    return idx, probs


# ----------------------------------------------------------------------------------------------
# transformer block with KV cache (hqvae/models/stage2/layers.py)
# ----------------------------------------------------------------------------------------------
def _linear(x: Tensor, P: Dict[str, Tensor], name: str) -> Tensor:
    return F.linear(x, P[name + ".weight"], P.get(name + ".bias"))


def block_sample(x: Tensor, P: Dict[str, Tensor], prefix: str, n_heads: int,
                 past_k: Optional[Tensor], past_v: Optional[Tensor], causal: bool,
                 rnd=_identity) -> Tuple[Tensor, Tensor, Tensor]:
    """`Block.sample` / `ParallelBlock.sample` (layers.py:324-328, 371-375) around
    `MultiHeadSelfAttention.forward(caching=True)` (layers.py:61-195).

    x [B,T,D] fp32 residual stream; past_k/past_v [B,nh,Tp,hs] or None.
    causal=True applies the tril mask of layers.py:106-111 / 118-123 among the T new tokens (only
    matters for T>1, i.e. the text prefill); causal=False is the ParallelBlock case where new tokens
    see everything (layers.py:127-152: the standard-attention mask at :130-137 only exists for
    past_kv=None with T=1, where it is the identity).
    Returns (x_out, k_new, v_new) with k_new/v_new [B,nh,T,hs]."""
    B, T, D = x.shape
    hs = D // n_heads
    h = rnd(F.layer_norm(x, (D,), P[prefix + ".ln1.weight"], P[prefix + ".ln1.bias"], 1e-5))    # :325
    q = rnd(_linear(h, P, prefix + ".attn.query")).view(B, T, n_heads, hs).transpose(1, 2)      # :73
    k = rnd(_linear(h, P, prefix + ".attn.key")).view(B, T, n_heads, hs).transpose(1, 2)        # :84
    v = rnd(_linear(h, P, prefix + ".attn.value")).view(B, T, n_heads, hs).transpose(1, 2)      # :85
    k_all = k if past_k is None else torch.cat([past_k, k], dim=2)                               # :93-96
    v_all = v if past_v is None else torch.cat([past_v, v], dim=2)
    Tp = k_all.shape[2] - T
    att = torch.matmul(q, k_all.transpose(-2, -1) * (1.0 / math.sqrt(hs)))                       # :102
    if causal and T > 1:
        mask = torch.ones(T, Tp + T, dtype=torch.bool)
        mask[:, Tp:] = torch.tril(torch.ones(T, T, dtype=torch.bool))
        att = att.masked_fill(~mask, float("-inf"))
    att = F.softmax(att, dim=-1)                                                                 # :183
    y = torch.matmul(att, v_all)                                                                 # :186
    y = rnd(y.transpose(1, 2).reshape(B, T, D))                                                  # :187
    x = x + _linear(y, P, prefix + ".attn.proj")                                                 # :190, :326
    h2 = rnd(F.layer_norm(x, (D,), P[prefix + ".ln2.weight"], P[prefix + ".ln2.bias"], 1e-5))
    m = rnd(F.gelu(_linear(h2, P, prefix + ".mlp.0")))                                           # :312-314 (erf)
    x = x + _linear(m, P, prefix + ".mlp.2")                                                     # :315, :327
    return x, k, v


# ----------------------------------------------------------------------------------------------
# one top position (hqvae/models/stage2/hierarchical_ar.py)
# ----------------------------------------------------------------------------------------------
def build_sos(P: Dict[str, Tensor], cfg: HQConfig, cond, num_candidates: int) -> Tensor:
    """sampling.py:183-192.  Extension over the reference: `cond` may also be an int64 [B] tensor of
    per-row class ids (the reference broadcasts one scalar class to the whole batch)."""
    if cfg.cond == "cls":
        if isinstance(cond, int):
            labels = torch.full((num_candidates,), cond, dtype=torch.long)
        else:
            labels = torch.as_tensor(cond, dtype=torch.long).reshape(-1)
        return P["sos.weight"][labels].unsqueeze(1)                                              # [B,1,D]
    if cfg.cond == "txt":
        ids = torch.as_tensor(cond, dtype=torch.long)
        return P["tok_emb_txt.weight"][ids] + P["pos_emb_txt.weight"][: cfg.idx_pred].unsqueeze(0)  # [B,64,D]
    return P["sos"].repeat(num_candidates, 1, 1)


def embed_stack(P: Dict[str, Tensor], code_top: Tensor, code_bot: Tensor, pos: int,
                cfg: Optional[HQConfig] = None) -> Tensor:
    """hierarchical_ar.py:506-507, 534-544 with `emb_blocks` empty (:100-113, n_layers_emb=1):
    mean over the 5 stack tokens of (embedding + pos_emb_emb[j]); the top token also gets
    pos_emb_top[pos].  code_top [B], code_bot [B,4] -> [B,1,D].
    position_embedding='2d' (:508-514): pos_emb_top_h[pos // H] + pos_emb_top_w[pos % H] with H = rows of the table
    (sqrt(ctx_len_img), NOT the 8 of the 8x8 grid - kept as the reference computes it).
    embedding_type='reduce' (:522-526): x = E_top[c_t] + pos + rearrange(E_bot[c_b], 'B (U L) K -> B U (K L)', U=1),
    i.e. element d of the bottom part is E_bot[c_b[d % 4]][d // 4] (K outer, L inner)."""
    if cfg is not None and cfg.position_embedding == "2d":
        H = P["pos_emb_top_h.weight"].shape[0]
        pos_emb = P["pos_emb_top_h.weight"][pos // H] + P["pos_emb_top_w.weight"][pos % H]
    else:
        pos_emb = P["pos_emb_top.weight"][pos]
    if cfg is not None and cfg.embedding_type == "reduce":
        e_top = P["tok_emb_top.weight"][code_top] + pos_emb                                      # [B,D]
        e_bot = P["tok_emb_bot.weight"][code_bot]                                                # [B,4,D/4]
        B = e_bot.shape[0]
        return (e_top + e_bot.permute(0, 2, 1).reshape(B, -1)).unsqueeze(1)                      # (K L) flattening
    e_top = P["tok_emb_top.weight"][code_top] + pos_emb                                          # [B,D]
    e_bot = P["tok_emb_bot.weight"][code_bot]                                                    # [B,4,D]
    h = torch.cat([e_top.unsqueeze(1), e_bot], dim=1) + P["pos_emb_emb.weight"].unsqueeze(0)     # [B,5,D]
    return h.mean(dim=1, keepdim=True)


class SpatialCache:
    """Pre-allocated replacement for the reference's list-of-presents `past`
    (sampling.py:227-231; re-concatenated at hierarchical_ar.py:554)."""

    def __init__(self, cfg: HQConfig, B: int, max_len: int):
        self.k = torch.zeros(cfg.n_layers, B, cfg.n_heads, max_len, cfg.head_dim)
        self.v = torch.zeros_like(self.k)
        self.len = 0


def spatial_step(P, cfg: HQConfig, x: Tensor, cache: SpatialCache, rnd=_identity) -> Tensor:
    """hierarchical_ar.py:482-563 after the input embedding: L x Block.sample then ln_f.
    x [B,T,D] (T=1, or ctx_len_txt for the text prefill) -> hs [B,T,D]."""
    T = x.shape[1]
    t0 = cache.len
    for l in range(cfg.n_layers):
        pk = cache.k[l, :, :, :t0] if t0 > 0 else None
        pv = cache.v[l, :, :, :t0] if t0 > 0 else None
        x, k, v = block_sample(x, P, f"blocks.{l}", cfg.n_heads, pk, pv, causal=True, rnd=rnd)
        cache.k[l, :, :, t0:t0 + T] = k
        cache.v[l, :, :, t0:t0 + T] = v
    cache.len = t0 + T
    D = cfg.embed_dim
    return F.layer_norm(x, (D,), P["ln_f.weight"], P["ln_f.bias"], 1e-5)                         # :561


def depth_pass0(P, cfg: HQConfig, hs_last: Tensor, rnd=_identity):
    """hierarchical_ar.py:682-695: y = hs + sos_depth -> Ld x ParallelBlock.sample (T=1, no past)
    -> head_top(ln_top(y)).  Returns (logits [B,V], list of (k,v) per depth layer)."""
    D = cfg.embed_dim
    y = hs_last + P["sos_depth"]
    kv = []
    for l in range(cfg.n_layers_depth):
        y, k, v = block_sample(y, P, f"depths.{l}", cfg.n_heads, None, None, causal=False, rnd=rnd)
        kv.append((k, v))
    y = rnd(F.layer_norm(y, (D,), P["ln_top.weight"], P["ln_top.bias"], 1e-5))
    return F.linear(y, P["head_top.weight"])[:, 0], kv


def depth_pass1(P, cfg: HQConfig, code_top: Tensor, kv0, rnd=_identity) -> Tensor:
    """hierarchical_ar.py:696-719: y_j = tok_emb_top_depth[c_top] + pos_emb_depth[j], j=0..3; each of
    the 4 queries attends to the pass-0 token and all 4 new tokens (no mask, layers.py:148-152);
    head_bot(ln_bot(y)).  Returns logits [B,4,V]."""
    D = cfg.embed_dim
    y = P["tok_emb_top_depth.weight"][code_top].unsqueeze(1) + P["pos_emb_depth.weight"][:4].unsqueeze(0)
    for l in range(cfg.n_layers_depth):
        y, _, _ = block_sample(y, P, f"depths.{l}", cfg.n_heads, kv0[l][0], kv0[l][1], causal=False, rnd=rnd)
    y = rnd(F.layer_norm(y, (D,), P["ln_bot.weight"], P["ln_bot.bias"], 1e-5))
    return F.linear(y, P["head_bot.weight"])


def depth_top2bot(P, cfg: HQConfig, hs_last: Tensor, draw, rnd=_identity):
    """`sampling_depth_baseline` / `sampling_step_depth_baseline` (hierarchical_ar.py:565-664), model_type='top2bot':
    five sequential single-token passes over the depth blocks with a growing cache.  Inputs: hs + sos_depth, then
    tok_emb_top_depth[c_top] + pos_emb_depth[0], then tok_emb_bot_depth[c_b(j-1)] + pos_emb_depth[j] (the position is the
    index of the code fed in, :630-637); outputs head_top(ln_top) for pass 0, head_bot(ln_bot) after.
    `draw(logits [B,V], slot)` returns the code [B] emitted for stack slot 0..4.  Returns (codes [B,5], logits list)."""
    D = cfg.embed_dim
    B = hs_last.shape[0]
    past = [(None, None)] * cfg.n_layers_depth
    codes, all_logits = [], []
    for c in range(5):
        if c == 0:
            y = hs_last + P["sos_depth"]
        elif c == 1:
            y = (P["tok_emb_top_depth.weight"][codes[0]] + P["pos_emb_depth.weight"][0]).unsqueeze(1)
        else:
            y = (P["tok_emb_bot_depth.weight"][codes[c - 1]] + P["pos_emb_depth.weight"][c - 1]).unsqueeze(1)
        new_past = []
        for l in range(cfg.n_layers_depth):
            y, k, v = block_sample(y, P, f"depths.{l}", cfg.n_heads, past[l][0], past[l][1], causal=True, rnd=rnd)
            pk, pv = past[l]
            new_past.append((k if pk is None else torch.cat([pk, k], 2), v if pv is None else torch.cat([pv, v], 2)))
        past = new_past
        if c == 0:
            y = rnd(F.layer_norm(y, (D,), P["ln_top.weight"], P["ln_top.bias"], 1e-5))
            logits = F.linear(y, P["head_top.weight"])[:, 0]
        else:
            y = rnd(F.layer_norm(y, (D,), P["ln_bot.weight"], P["ln_bot.bias"], 1e-5))
            logits = F.linear(y, P["head_bot.weight"])[:, 0]
        all_logits.append(logits)
        codes.append(draw(logits, c))
    return torch.stack(codes, 1), all_logits


def depth_bidirectional(P, cfg: HQConfig, hs_last: Tensor, rnd=_identity):
    """`sampling_step_depth_bidirectional` (hierarchical_ar.py:791-826), model_type='bidirectional': ONE pass over the five
    tokens [hs + sos_depth, pos_emb_depth[0..3]] with unmasked attention among them (Block(causal_attn=False));
    logits_top = head_top(ln_top(y[:, 0])), logits_bot = head_bot(ln_bot(y[:, 1:])).  Returns (top [B,V], bot [B,4,V])."""
    D = cfg.embed_dim
    B = hs_last.shape[0]
    y = torch.cat([hs_last + P["sos_depth"], P["pos_emb_depth.weight"][:4].unsqueeze(0).repeat(B, 1, 1)], dim=1)
    for l in range(cfg.n_layers_depth):
        y, _, _ = block_sample(y, P, f"depths.{l}", cfg.n_heads, None, None, causal=False, rnd=rnd)
    yt = rnd(F.layer_norm(y[:, 0:1], (D,), P["ln_top.weight"], P["ln_top.bias"], 1e-5))
    yb = rnd(F.layer_norm(y[:, 1:], (D,), P["ln_bot.weight"], P["ln_bot.bias"], 1e-5))
    return F.linear(yt, P["head_top.weight"])[:, 0], F.linear(yb, P["head_bot.weight"])


def _as_pair(softmax_temperature) -> Tuple[float, float]:
    """The reference indexes `softmax_temperatures[0/1]` (hierarchical_ar.py:763, 779); the shipped
    measure_throughput_txt passes a bare float - accept both."""
    if isinstance(softmax_temperature, (int, float)):
        return float(softmax_temperature), float(softmax_temperature)
    return float(softmax_temperature[0]), float(softmax_temperature[1])


@torch.no_grad()
def sample(params: Dict[str, Tensor], cfg: HQConfig, cond, num_candidates: int,
           top_k_top: Optional[int] = None, top_p_top: Optional[float] = None,
           top_k_bot: Optional[int] = None, top_p_bot: Optional[float] = None,
           softmax_temperature=(1.0, 1.0), max_seq_len: int = 64,
           given_top_code: Optional[Tensor] = None, given_bot_code: Optional[Tensor] = None,
           emulate: Optional[str] = None, generator: Optional[torch.Generator] = None,
           return_logits: bool = False):
    """`sampling_ihqgpt` (sampling.py:164-237) + `iHQGPT.sampling_step` (hierarchical_ar.py:428-480)
    + `sampling_depth_parallel` (:721-789).

    Returns codes_top [B,max_seq_len] int64, codes_bot [B,max_seq_len,4] int64
    (+ logits [B,max_seq_len,5,V] fp32 when return_logits).  Draw order per position: top, b0..b3,
    one `torch.multinomial` call each over the whole batch - identical to the reference, so with the
    same torch seed and fp32 the sequences are identical too.
    `given_top_code` [B,S] / `given_bot_code` [B,S,4] force the emitted codes (teacher forcing)."""
    rnd = _round_bf16 if emulate == "bf16" else _identity
    P = round_params(params, emulate)
    T_top, T_bot = _as_pair(softmax_temperature)
    B = num_candidates
    sos = build_sos(P, cfg, cond, B)
    assert sos.shape[0] == B
    cache = SpatialCache(cfg, B, cfg.idx_pred + max_seq_len)
    codes_top = torch.zeros(B, max_seq_len, dtype=torch.long)
    codes_bot = torch.zeros(B, max_seq_len, 4, dtype=torch.long)
    V = max(cfg.vocab_top, cfg.vocab_bot)
    all_logits = torch.zeros(B, max_seq_len, 5, V) if return_logits else None

    for cnt in range(max_seq_len):
        if cnt == 0:
            x = sos                                                                              # :493-499
        else:
            x = embed_stack(P, codes_top[:, cnt - 1], codes_bot[:, cnt - 1], cnt - 1, cfg)       # :506-544
        hs = spatial_step(P, cfg, x, cache, rnd)
        hs_last = hs[:, -1:, :]                                                                  # :684-685
        if cfg.model_type == "top2bot":
            def draw(logits, slot):                                                              # :639-661
                if slot == 0 and given_top_code is not None:
                    return given_top_code[:, cnt]
                if slot > 0 and given_bot_code is not None:
                    return given_bot_code[:, cnt, slot - 1]
                T, k, p = (T_top, top_k_top, top_p_top) if slot == 0 else (T_bot, top_k_bot, top_p_bot)
                return draw_token(logits, T, k, p, generator)[0][:, 0]
            codes, lgs = depth_top2bot(P, cfg, hs_last, draw, rnd)
            codes_top[:, cnt] = codes[:, 0]
            codes_bot[:, cnt] = codes[:, 1:]
            if return_logits:
                all_logits[:, cnt, 0, : cfg.vocab_top] = lgs[0]
                for j in range(4):
                    all_logits[:, cnt, 1 + j, : cfg.vocab_bot] = lgs[1 + j]
            continue
        if cfg.model_type == "bidirectional":
            # :846-871: every one of the five tokens is drawn with the BOTTOM filters and softmax_temperatures[0]
            lt, lb = depth_bidirectional(P, cfg, hs_last, rnd)
            if return_logits:
                all_logits[:, cnt, 0, : cfg.vocab_top] = lt
                all_logits[:, cnt, 1:, : cfg.vocab_bot] = lb
            if given_top_code is not None:
                codes_top[:, cnt] = given_top_code[:, cnt]
            else:
                codes_top[:, cnt] = draw_token(lt, T_top, top_k_bot, top_p_bot, generator)[0][:, 0]
            for j in range(4):
                if given_bot_code is not None:
                    codes_bot[:, cnt, j] = given_bot_code[:, cnt, j]
                else:
                    codes_bot[:, cnt, j] = draw_token(lb[:, j], T_top, top_k_bot, top_p_bot, generator)[0][:, 0]
            continue
        logits_top, kv0 = depth_pass0(P, cfg, hs_last, rnd)
        if given_top_code is None:
            c_top, _ = draw_token(logits_top, T_top, top_k_top, top_p_top, generator)            # :763-769
            c_top = c_top[:, 0]
        else:
            c_top = given_top_code[:, cnt]                                                       # :771-772
        logits_bot = depth_pass1(P, cfg, c_top, kv0, rnd)
        if return_logits:
            all_logits[:, cnt, 0, : cfg.vocab_top] = logits_top
            all_logits[:, cnt, 1:, : cfg.vocab_bot] = logits_bot
        codes_top[:, cnt] = c_top
        for j in range(4):                                                                       # :778-785
            if given_bot_code is None:
                c, _ = draw_token(logits_bot[:, j], T_bot, top_k_bot, top_p_bot, generator)
                codes_bot[:, cnt, j] = c[:, 0]
            else:
                codes_bot[:, cnt, j] = given_bot_code[:, cnt, j]
    if return_logits:
        return codes_top, codes_bot, all_logits
    return codes_top, codes_bot


@torch.no_grad()
def step_logits(params, cfg: HQConfig, cond, codes_top: Tensor, codes_bot: Tensor,
                emulate: Optional[str] = None) -> Tensor:
    """Teacher-forced logits [B,S,5,V] for given code grids (the incremental path, not the training
    forward): what `sampling_step` would have seen had it emitted exactly these codes."""
    B, S = codes_top.shape
    _, _, lg = sample(params, cfg, cond, B, max_seq_len=S, given_top_code=codes_top,
                      given_bot_code=codes_bot, emulate=emulate, return_logits=True)
    return lg


def codes_to_grids(codes_top: Tensor, codes_bot: Tensor, H: int = 8) -> Tuple[Tensor, Tensor]:
    """HQ-VAE code layout consumed by `stage1.decode_code` (sampling_hqmodel.py:119-120):
    'B (H W) -> B H W' and 'B (H W) (kerH kerW) -> B (H kerH) (W kerW)' with kerH=kerW=2."""
    B = codes_top.shape[0]
    W = codes_top.shape[1] // H
    top = codes_top.view(B, H, W)
    bot = codes_bot.view(B, H, W, 2, 2).permute(0, 1, 3, 2, 4).reshape(B, 2 * H, 2 * W)
    return top, bot
