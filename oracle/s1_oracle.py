"""TEST INFRASTRUCTURE ONLY - CPU oracle for stage-1 decoding of sampled code grids.

Functional fp32 restatement (torch CPU) of `SimRQGAN2Generator.decode_code` for the shipped HQ-VAE configuration
(`decoding_type: concat`, `upsample: pixelshuffle`, EMA quantizers):

    hqvae/models/stage1/generator.py:312-367      decode_code / decode
    hqvae/models/stage1/modules/quantizer.py:179-186   get_codebook_entry
    hqvae/models/stage1/modules/layers.py:12-22, 36-53, 77-186, 300-410   swish, GroupNorm(32, eps 1e-6), Upsample,
                                                                            ResnetBlock, AttnBlock, Decoder

It exists to CHECK the CUDA path (hqtransformer_b200/stage1.py -> libhqgraft hq_s1_*); it is never the product and never
a fallback.  Pinned by running the unmodified reference module (oracle/ref_shim.py: build_reference_stage1) in the build
container: tests/golden/s1_*.npz (oracle/make_golden.py --stage1) and a live cross-check in tests/test_oracle_golden.py.
`emulate='bf16'` rounds every convolution / attention input and the conv weights to bf16 (fp32 accumulation, fp32
residual stream and GroupNorm), which is where the CUDA path rounds.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, asdict, field
from typing import Dict, List, Optional, Tuple

import torch
from torch.nn import functional as F

Tensor = torch.Tensor


@dataclass
class S1Config:
    """`stage1` section of the stage-2 YAMLs (e.g. configs/master/stage2/imagenet/hqtransformer-embtrans1-soft1-layer12-top8x8.yaml:5-27)."""
    embed_dim: int = 256
    n_embed: int = 8192
    z_channels: int = 256
    resolution: int = 256
    ch: int = 128
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    attn_resolutions: Tuple[int, ...] = (16,)
    out_ch: int = 3

    @property
    def n_levels(self) -> int:
        return len(self.ch_mult)

    @property
    def latent_res(self) -> int:          # use_init_downsample: True -> resolution / 2^num_resolutions (layers.py:331-332)
        return self.resolution // 2 ** self.n_levels

    def to_dict(self):
        d = asdict(self)
        d["ch_mult"] = list(self.ch_mult)
        d["attn_resolutions"] = list(self.attn_resolutions)
        return d

    @staticmethod
    def from_dict(d):
        d = dict(d)
        d["ch_mult"] = tuple(d["ch_mult"])
        d["attn_resolutions"] = tuple(d["attn_resolutions"])
        return S1Config(**d)


IMAGENET_S1 = S1Config()
# every conv input a multiple of 64 channels (the CUDA path's k-block), 4 x 4 top / 8 x 8 bottom grids, 64 x 64 pixels
TINY_S1 = S1Config(embed_dim=64, n_embed=128, z_channels=64, resolution=64, ch=64, ch_mult=(1, 2, 2, 2), num_res_blocks=1,
                   attn_resolutions=(4,))


def _res_shapes(prefix: str, cin: int, cout: int, s):
    s[f"{prefix}.norm1.weight"] = (cin,)
    s[f"{prefix}.norm1.bias"] = (cin,)
    s[f"{prefix}.conv1.weight"] = (cout, cin, 3, 3)
    s[f"{prefix}.conv1.bias"] = (cout,)
    s[f"{prefix}.norm2.weight"] = (cout,)
    s[f"{prefix}.norm2.bias"] = (cout,)
    s[f"{prefix}.conv2.weight"] = (cout, cout, 3, 3)
    s[f"{prefix}.conv2.bias"] = (cout,)
    if cin != cout:                                                     # layers.py:103-116 (nin_shortcut, 1 x 1)
        s[f"{prefix}.nin_shortcut.weight"] = (cout, cin, 1, 1)
        s[f"{prefix}.nin_shortcut.bias"] = (cout,)


def _attn_shapes(prefix: str, c: int, s):
    s[f"{prefix}.norm.weight"] = (c,)
    s[f"{prefix}.norm.bias"] = (c,)
    for nm in ("q", "k", "v", "proj_out"):
        s[f"{prefix}.{nm}.weight"] = (c, c, 1, 1)
        s[f"{prefix}.{nm}.bias"] = (c,)


def decoder_plan(cfg: S1Config):
    """Structure of `Decoder.__init__` (layers.py:300-383) for use_init_downsample / use_mid_block / use_attn = True:
    list of (level, [(block_in, block_out, has_attn), ...], upsample_channels) from the lowest resolution up."""
    block_in = cfg.ch * cfg.ch_mult[-1]
    res = cfg.latent_res
    plan = []
    for lvl in reversed(range(cfg.n_levels)):
        block_out = cfg.ch * cfg.ch_mult[lvl]
        blocks = []
        for _ in range(cfg.num_res_blocks + 1):
            blocks.append((block_in, block_out, res in cfg.attn_resolutions))
            block_in = block_out
        plan.append((lvl, blocks, block_in, res))
        res *= 2
    return plan


def param_shapes(cfg: S1Config) -> "OrderedDict[str, Tuple[int, ...]]":
    """The state_dict keys of the reference generator that `decode_code` reads (generator.py:243-250, layers.py:300-383)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["quantize_t.embedding"] = (cfg.n_embed, cfg.embed_dim * 4)       # pixelshuffle: top entries carry 2 x 2 bottom cells
    s["quantize_b.embedding"] = (cfg.n_embed, cfg.embed_dim)
    s["post_quant_conv_b.weight"] = (cfg.z_channels, cfg.embed_dim * 2, 1, 1)
    s["post_quant_conv_b.bias"] = (cfg.z_channels,)
    top = cfg.ch * cfg.ch_mult[-1]
    s["decoder.conv_in.weight"] = (top, cfg.z_channels, 3, 3)
    s["decoder.conv_in.bias"] = (top,)
    _res_shapes("decoder.mid.block_1", top, top, s)
    _attn_shapes("decoder.mid.attn_1", top, s)
    _res_shapes("decoder.mid.block_2", top, top, s)
    for lvl, blocks, up_ch, _ in decoder_plan(cfg):
        for j, (cin, cout, has_attn) in enumerate(blocks):
            _res_shapes(f"decoder.up.{lvl}.block.{j}", cin, cout, s)
            if has_attn:
                _attn_shapes(f"decoder.up.{lvl}.attn.{j}", cout, s)
        s[f"decoder.up.{lvl}.upsample.conv.weight"] = (up_ch, up_ch, 3, 3)
        s[f"decoder.up.{lvl}.upsample.conv.bias"] = (up_ch,)
    last = cfg.ch * cfg.ch_mult[0]
    s["decoder.norm_out.weight"] = (last,)
    s["decoder.norm_out.bias"] = (last,)
    s["decoder.conv_out.weight"] = (cfg.out_ch, last, 3, 3)
    s["decoder.conv_out.bias"] = (cfg.out_ch,)
    return s


def make_params(cfg: S1Config, seed: int = 0) -> "OrderedDict[str, Tensor]":
    """Deterministic synthetic weights with trained-checkpoint-like statistics: conv weights ~ N(0, 1/fan_in) (activations
    stay O(1) through ~40 layers), small random biases, GroupNorm affines around (1, 0), codebooks ~ N(0, 1)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        leaf = name.split(".")[-1]
        if "norm" in name:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g) if leaf == "weight" else 0.05 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        elif leaf == "embedding":
            t = torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        out[name] = t
    return out


def _rb(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _swish(x):                                                          # layers.py:12-14
    return x * torch.sigmoid(x)


def _gn(x, P, name):                                                    # layers.py:17-21
    return F.group_norm(x, 32, P[name + ".weight"], P[name + ".bias"], eps=1e-6)


def _conv(x, P, name, pad, rnd, rw):
    return F.conv2d(rnd(x), rw(P[name + ".weight"]), P[name + ".bias"], padding=pad)


def _resblock(x, P, prefix, rnd, rw):                                   # layers.py:118-135
    h = _conv(_swish(_gn(x, P, prefix + ".norm1")), P, prefix + ".conv1", 1, rnd, rw)
    h = _conv(_swish(_gn(h, P, prefix + ".norm2")), P, prefix + ".conv2", 1, rnd, rw)
    if prefix + ".nin_shortcut.weight" in P:
        x = _conv(x, P, prefix + ".nin_shortcut", 0, rnd, rw)
    return x + h


def _attnblock(x, P, prefix, rnd, rw):                                  # layers.py:163-186
    h = _gn(x, P, prefix + ".norm")
    q = rnd(_conv(h, P, prefix + ".q", 0, rnd, rw))
    k = rnd(_conv(h, P, prefix + ".k", 0, rnd, rw))
    v = rnd(_conv(h, P, prefix + ".v", 0, rnd, rw))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, P, prefix + ".proj_out", 0, rnd, rw)


@torch.no_grad()
def decode_code(P: Dict[str, Tensor], cfg: S1Config, code_t: Tensor, code_b: Tensor, emulate: Optional[str] = None) -> Tensor:
    """generator.py:323-367 with both grids given: code_t [B, h, w], code_b [B, 2h, 2w] (int64) -> pixels [B, 3, R, R]."""
    rnd = _rb if emulate == "bf16" else (lambda t: t)
    rw = rnd
    quant_t = P["quantize_t.embedding"][code_t].permute(0, 3, 1, 2)       # quantizer.py:182-186; [B, 4E, h, w]
    quant_b = P["quantize_b.embedding"][code_b].permute(0, 3, 1, 2)       # [B, E, 2h, 2w]
    quant = torch.cat([F.pixel_shuffle(quant_t, 2), quant_b], dim=1)      # generator.py:316-318
    h = _conv(quant, P, "post_quant_conv_b", 0, rnd, rw)                  # :319
    h = _conv(h, P, "decoder.conv_in", 1, rnd, rw)                        # layers.py:389
    h = _resblock(h, P, "decoder.mid.block_1", rnd, rw)
    h = _attnblock(h, P, "decoder.mid.attn_1", rnd, rw)
    h = _resblock(h, P, "decoder.mid.block_2", rnd, rw)
    for lvl, blocks, _, _ in decoder_plan(cfg):                           # :398-404
        for j, (_, _, has_attn) in enumerate(blocks):
            h = _resblock(h, P, f"decoder.up.{lvl}.block.{j}", rnd, rw)
            if has_attn:
                h = _attnblock(h, P, f"decoder.up.{lvl}.attn.{j}", rnd, rw)
        h = F.interpolate(h, scale_factor=2.0, mode="nearest")            # Upsample, :49-52
        h = _conv(h, P, f"decoder.up.{lvl}.upsample.conv", 1, rnd, rw)
    h = _swish(_gn(h, P, "decoder.norm_out"))                             # :406-408
    return _conv(h, P, "decoder.conv_out", 1, rnd, rw)
