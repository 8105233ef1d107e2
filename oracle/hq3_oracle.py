"""TEST INFRASTRUCTURE ONLY - CPU oracle for the 3-level `HQTransformer` sampling loop.

Functional fp32 restatement (torch CPU) of the reference's multi-level sampler for `decoding_type: parallel-add`,
`embedding_type: transformer1` (the shipped `*-level3.yaml` configs):

    hqvae/utils/sampling.py:240-307                      sampling_hqtransformer (outer loop over the top positions)
    hqvae/models/stage2/hqtransformer.py:409-439         sampling_step
    hqvae/models/stage2/hqtransformer.py:441-498         sampling_step_spatial (21-token stack embedding, mean)
    hqvae/models/stage2/hqtransformer.py:500-571         sampling_step_hierarchy_parallel (three depth passes: 1, 4, 16 tokens)
    hqvae/models/stage2/hqtransformer.py:573-635         sampling_hierarchy_parallel (21 categorical draws)
    hqvae/models/stage2/layers.py:154-178                ParallelBlock mask for code_level == 3, 'parallel'

It exists to CHECK the CUDA path; it is never the product and never a fallback.  Pinned by goldens made with the unmodified
reference (oracle/make_golden.py --level3) and by a live cross-check (tests/test_oracle_golden.py).
Stack layout per top position: slot 0 = top code, slots 1..4 = middle codes (raster of the 2x2 cell), slots 5..20 = bottom
codes (raster of the 4x4 cell).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, asdict
from typing import Dict, Optional, Sequence, Tuple

import torch
from torch.nn import functional as F

from oracle.hq_oracle import (_block_shapes, _identity, _round_bf16, block_sample, draw_token, round_params, SpatialCache,
                              spatial_step)

Tensor = torch.Tensor
N_STACK = 21


@dataclass
class HQ3Config:
    """`HQTransformer(vocab_sizes, ..., decoding_type='parallel-add')` (hqtransformer.py:168-216)."""
    embed_dim: int = 1536
    n_heads: int = 24
    n_layers: int = 12
    n_layers_depth: int = 4
    vocab_sizes: Tuple[int, int, int] = (8192, 8192, 8192)
    n_classes: int = 1000
    ctx_len_img: int = 256
    cond: str = "cls"                # 'cls' | 'uncond'

    @property
    def head_dim(self):
        return self.embed_dim // self.n_heads

    idx_pred = 0
    ctx_len_txt = 0

    def to_dict(self):
        d = asdict(self)
        d["vocab_sizes"] = list(self.vocab_sizes)
        return d

    @staticmethod
    def from_dict(d):
        d = dict(d)
        d["vocab_sizes"] = tuple(d["vocab_sizes"])
        return HQ3Config(**d)


TINY3 = HQ3Config(embed_dim=128, n_heads=2, n_layers=2, n_layers_depth=2, vocab_sizes=(256, 192, 320), n_classes=10,
                  ctx_len_img=64)
SMALL3 = HQ3Config(embed_dim=256, n_heads=4, n_layers=3, n_layers_depth=3, vocab_sizes=(512, 512, 1024), n_classes=10,
                   ctx_len_img=64)
IMAGENET_L12_LEVEL3 = HQ3Config()


def param_shapes(cfg: HQ3Config) -> "OrderedDict[str, Tuple[int, ...]]":
    """state_dict keys of the reference HQTransformer (hqtransformer.py:24-166) for parallel-add / transformer1 / 1d."""
    D = cfg.embed_dim
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for i, v in enumerate(cfg.vocab_sizes):
        s[f"tok_emb_levels.{i}.weight"] = (v, D)                    # :37-40
    s["pos_emb_emb.weight"] = (N_STACK, D)                         # :41
    if cfg.cond == "cls":
        s["sos.weight"] = (cfg.n_classes, D)                       # :61
    else:
        s["sos"] = (1, 1, D)                                       # :74
    s["pos_emb_top.weight"] = (cfg.ctx_len_img, D)                 # :80
    for i in range(cfg.n_layers):
        s.update(_block_shapes(f"blocks.{i}", D))
    s["ln_f.weight"] = (D,)
    s["ln_f.bias"] = (D,)
    s["sos_depth"] = (1, 1, D)                                     # :102
    for i, v in enumerate(cfg.vocab_sizes):
        s[f"tok_emb_depth_levels.{i}.weight"] = (v, D)             # :106-117 (level 2 is never read when sampling)
    s["pos_emb_depths.0.weight"] = (4, D)                          # :125-128
    s["pos_emb_depths.1.weight"] = (16, D)
    for i in range(cfg.n_layers_depth):
        s.update(_block_shapes(f"depths.{i}", D))
    for i in range(3):                                             # :161-166
        s[f"ln_levels.{i}.weight"] = (D,)
        s[f"ln_levels.{i}.bias"] = (D,)
    for i, v in enumerate(cfg.vocab_sizes):
        s[f"head_levels.{i}.weight"] = (v, D)
    # module registration order of the reference interleaves ln / head per level; order is irrelevant for loading
    return s


def make_params(cfg: HQ3Config, seed: int = 0, init: str = "rich") -> "OrderedDict[str, Tensor]":
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        leaf = name.split(".")[-1]
        is_ln = (".ln" in name or name.startswith("ln_")) and "head" not in name
        if name in ("sos", "sos_depth"):
            t = torch.randn(shape, generator=g)
        elif is_ln:
            if init == "rich":
                t = (1.0 + 0.1 * torch.randn(shape, generator=g)) if leaf == "weight" else 0.05 * torch.randn(shape, generator=g)
            else:
                t = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        elif leaf == "bias":
            t = 0.02 * torch.randn(shape, generator=g) if init == "rich" else torch.zeros(shape)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t
    return out


def _round_params3(P, emulate):
    if emulate is None:
        return P
    out = round_params(P, emulate)
    for k in list(out):
        if k.startswith("head_levels."):
            out[k] = _round_bf16(P[k])
    return out


def embed_stack3(P, ct: Tensor, cm: Tensor, cb: Tensor, pos: int) -> Tensor:
    """hqtransformer.py:453-487: mean over the 21 stack tokens of (embedding + pos_emb_emb[j]); the top token also gets
    pos_emb_top[pos].  ct [B], cm [B,4], cb [B,16] -> [B,1,D]."""
    e0 = (P["tok_emb_levels.0.weight"][ct] + P["pos_emb_top.weight"][pos]).unsqueeze(1)
    e1 = P["tok_emb_levels.1.weight"][cm]
    e2 = P["tok_emb_levels.2.weight"][cb]
    h = torch.cat([e0, e1, e2], dim=1) + P["pos_emb_emb.weight"].unsqueeze(0)
    return h.mean(dim=1, keepdim=True)


def _ln_head(P, y, level, rnd):
    D = y.shape[-1]
    y = rnd(F.layer_norm(y, (D,), P[f"ln_levels.{level}.weight"], P[f"ln_levels.{level}.bias"], 1e-5))
    return F.linear(y, P[f"head_levels.{level}.weight"])


def depth_passes(P, cfg: HQ3Config, hs_last: Tensor, draw, rnd=_identity):
    """hqtransformer.py:500-635 for 'parallel-add'.  `draw(logits [B,V], slot 0..20)` -> code [B].
    Returns (codes [B,21], logits list of 21 [B,V_level])."""
    B = hs_last.shape[0]
    Ld = cfg.n_layers_depth
    # pass 0: the top code (:515-519, mask row 0)
    y = hs_last + P["sos_depth"]
    kv = []
    for l in range(Ld):
        y, k, v = block_sample(y, P, f"depths.{l}", cfg.n_heads, None, None, causal=False, rnd=rnd)
        kv.append((k, v))
    lg0 = _ln_head(P, y, 0, rnd)[:, 0]
    logits = [lg0]
    codes = [draw(lg0, 0)]
    # pass 1: four middle codes (:521-531; mask rows 1..4 see columns 0..4)
    y = P["tok_emb_depth_levels.0.weight"][codes[0]].unsqueeze(1) + P["pos_emb_depths.0.weight"].unsqueeze(0)
    kv1 = []
    for l in range(Ld):
        y, k, v = block_sample(y, P, f"depths.{l}", cfg.n_heads, kv[l][0], kv[l][1], causal=False, rnd=rnd)
        kv1.append((torch.cat([kv[l][0], k], 2), torch.cat([kv[l][1], v], 2)))
    lg1 = _ln_head(P, y, 1, rnd)
    for j in range(4):
        logits.append(lg1[:, j])
        codes.append(draw(lg1[:, j], 1 + j))
    # pass 2: sixteen bottom codes in raster order of the 4x4 cell (:521-548): token t = (row, col) takes the embedding of
    # its middle parent (row // 2, col // 2), the position embedding of t, and ('add') the top-code embedding
    mid = torch.stack(codes[1:5], 1)                                           # [B,4], raster of the 2x2 cell
    t = torch.arange(16)
    parent = (t // 4 // 2) * 2 + (t % 4) // 2
    y = (P["tok_emb_depth_levels.1.weight"][mid[:, parent]] + P["pos_emb_depths.1.weight"].unsqueeze(0)
         + P["tok_emb_depth_levels.0.weight"][codes[0]].unsqueeze(1))
    for l in range(Ld):
        y, _, _ = block_sample(y, P, f"depths.{l}", cfg.n_heads, kv1[l][0], kv1[l][1], causal=False, rnd=rnd)
    lg2 = _ln_head(P, y, 2, rnd)
    for j in range(16):
        logits.append(lg2[:, j])
        codes.append(draw(lg2[:, j], 5 + j))
    return torch.stack(codes, 1), logits


@torch.no_grad()
def sample(params: Dict[str, Tensor], cfg: HQ3Config, cond, num_candidates: int, top_k: Sequence[Optional[int]] = (None, None, None),
           top_p: Sequence[Optional[float]] = (None, None, None), softmax_temperature: Sequence[float] = (1.0, 1.0, 1.0),
           max_seq_len: int = 64, given: Optional[Tensor] = None, emulate: Optional[str] = None,
           generator: Optional[torch.Generator] = None, return_logits: bool = False):
    """`sampling_hqtransformer` (sampling.py:240-307).  Returns [codes_top [B,S], codes_mid [B,S,4], codes_bot [B,S,16]]
    (+ logits [B,S,21,Vmax]).  `given` [B,S,21] forces every emitted code (teacher forcing)."""
    rnd = _round_bf16 if emulate == "bf16" else _identity
    P = _round_params3(params, emulate)
    B = num_candidates
    if cfg.cond == "cls":
        labels = torch.full((B,), cond, dtype=torch.long) if isinstance(cond, int) else torch.as_tensor(cond, dtype=torch.long).reshape(-1)
        sos = P["sos.weight"][labels].unsqueeze(1)
    else:
        sos = P["sos"].repeat(B, 1, 1)
    cache = SpatialCache(cfg, B, max_seq_len)
    codes = torch.zeros(B, max_seq_len, N_STACK, dtype=torch.long)
    Vmax = max(cfg.vocab_sizes)
    all_logits = torch.zeros(B, max_seq_len, N_STACK, Vmax) if return_logits else None
    level_of = [0] + [1] * 4 + [2] * 16
    for cnt in range(max_seq_len):
        if cnt == 0:
            x = sos
        else:
            c = codes[:, cnt - 1]
            x = embed_stack3(P, c[:, 0], c[:, 1:5], c[:, 5:], cnt - 1)
        hs = spatial_step(P, cfg, x, cache, rnd)

        def draw(logits, slot):
            if given is not None:
                return given[:, cnt, slot]
            lv = level_of[slot]
            return draw_token(logits, softmax_temperature[lv], top_k[lv], top_p[lv], generator)[0][:, 0]
        cc, lgs = depth_passes(P, cfg, hs[:, -1:], draw, rnd)
        codes[:, cnt] = cc
        if return_logits:
            for j, lg in enumerate(lgs):
                all_logits[:, cnt, j, : lg.shape[-1]] = lg
    out = [codes[:, :, 0], codes[:, :, 1:5], codes[:, :, 5:]]
    return (out, all_logits) if return_logits else out
