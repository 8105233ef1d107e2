"""-m gpu: stage-1 decode of code grids through the C ABI (hq_s1_*): implicit-GEMM tcgen05 convolutions,
GroupNorm, attention, against the reference-made golden, the fp32 oracle and the bf16-emulating oracle."""
import numpy as np
import pytest
import torch

from oracle import s1_oracle as S1
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _decoder(cfg, P, max_batch):
    import hqtransformer_b200 as H
    dec = H.HQVAEDecoder(embed_dim=cfg.embed_dim, n_embed=cfg.n_embed, z_channels=cfg.z_channels, resolution=cfg.resolution,
                         ch=cfg.ch, ch_mult=cfg.ch_mult, num_res_blocks=cfg.num_res_blocks,
                         attn_resolutions=cfg.attn_resolutions, out_ch=cfg.out_ch, max_batch=max_batch)
    assert list(dec.param_shapes().items()) == list(S1.param_shapes(cfg).items())
    dec.load_state_dict(P, strict=True)
    return dec


def test_stage1_decode_matches_reference_golden():
    """Pixels of the unmodified `SimRQGAN2Generator.decode_code` (CPU fp32).  The CUDA path feeds bf16 into every
    convolution (fp32 accumulation, fp32 residual stream / GroupNorm): bar max-abs <= 8e-2, mean-abs <= 1e-2 on pixels of
    mean magnitude ~0.45 against the fp32 reference, and max-abs <= 6e-2, mean <= 6e-3 against the oracle that rounds at the
    same points (not tighter than the fp32 bar: a GroupNorm output near a bf16 rounding boundary flips with the summation
    order, and ~40 layers follow)."""
    g, meta = load_golden("s1_tiny_decode.npz")
    cfg = S1.S1Config.from_dict(meta["config"])
    P = S1.make_params(cfg, seed=meta["seed"])
    ct, cb = torch.from_numpy(g["code_t"]), torch.from_numpy(g["code_b"])
    dec = _decoder(cfg, P, max_batch=4)
    got = dec.decode_code(ct, cb).cpu()
    want = torch.from_numpy(g["pixels"])
    assert tuple(got.shape) == tuple(want.shape) and torch.isfinite(got).all()
    err = (got - want).abs()
    assert float(err.max()) <= 8e-2 and float(err.mean()) <= 1e-2, (float(err.max()), float(err.mean()))
    emu = S1.decode_code(P, cfg, ct, cb, emulate="bf16")
    e2 = (got - emu).abs()
    assert float(e2.max()) <= 6e-2 and float(e2.mean()) <= 6e-3, (float(e2.max()), float(e2.mean()))


@pytest.mark.parametrize("B,max_batch", [(1, 1), (5, 2), (7, 8)])
def test_stage1_decode_batches_and_chunks(B, max_batch):
    """Any batch (chunked above max_batch), deterministic, an image does not depend on its neighbours in the batch."""
    cfg = S1.TINY_S1
    P = S1.make_params(cfg, seed=3)
    g = torch.Generator().manual_seed(B)
    h = cfg.latent_res // 2
    ct = torch.randint(0, cfg.n_embed, (B, h, h), generator=g)
    cb = torch.randint(0, cfg.n_embed, (B, 2 * h, 2 * h), generator=g)
    dec = _decoder(cfg, P, max_batch=max_batch)
    a = dec.decode_code(ct, cb)
    b = dec.decode_code(ct, cb)
    assert torch.equal(a, b)
    single = dec.decode_code(ct[B - 1:], cb[B - 1:])
    assert torch.equal(single[0], a[B - 1])
    emu = S1.decode_code(P, cfg, ct, cb, emulate="bf16")
    e = (a.cpu() - emu).abs()
    assert float(e.max()) <= 6e-2 and float(e.mean()) <= 6e-3, (float(e.max()), float(e.mean()))


def test_stage1_full_size_decode_vs_oracle():
    """The shipped HQ-VAE decoder (256 x 256 pixels, ch 128, ch_mult [1,2,4,4], attention at 16 x 16; 8 x 8 + 16 x 16 grids)
    on two images against the oracle (fp32 and bf16-emulating)."""
    cfg = S1.IMAGENET_S1
    P = S1.make_params(cfg, seed=2)
    g = torch.Generator().manual_seed(5)
    ct = torch.randint(0, cfg.n_embed, (2, 8, 8), generator=g)
    cb = torch.randint(0, cfg.n_embed, (2, 16, 16), generator=g)
    dec = _decoder(cfg, P, max_batch=2)
    got = dec.decode_code(ct, cb).cpu()
    want = S1.decode_code(P, cfg, ct, cb)
    emu = S1.decode_code(P, cfg, ct, cb, emulate="bf16")
    e1, e2 = (got - want).abs(), (got - emu).abs()
    assert float(e1.max()) <= 1e-1 and float(e1.mean()) <= 1e-2, (float(e1.max()), float(e1.mean()))
    assert float(e2.max()) <= 1e-1 and float(e2.mean()) <= 6e-3, (float(e2.max()), float(e2.mean()))


def test_stage1_errors_are_loud():
    import hqtransformer_b200 as H
    cfg = S1.TINY_S1
    P = S1.make_params(cfg, seed=3)
    dec = _decoder(cfg, P, max_batch=2)
    h = cfg.latent_res // 2
    with pytest.raises(IndexError):
        dec.decode_code(torch.full((1, h, h), cfg.n_embed), torch.zeros(1, 2 * h, 2 * h, dtype=torch.long))
    with pytest.raises(ValueError):
        dec.decode_code(torch.zeros(1, h + 1, h, dtype=torch.long), torch.zeros(1, 2 * h, 2 * h, dtype=torch.long))
    fresh = H.HQVAEDecoder(embed_dim=cfg.embed_dim, n_embed=cfg.n_embed, z_channels=cfg.z_channels, resolution=cfg.resolution,
                           ch=cfg.ch, ch_mult=cfg.ch_mult, num_res_blocks=cfg.num_res_blocks,
                           attn_resolutions=cfg.attn_resolutions, max_batch=1)
    with pytest.raises(H.HQError):
        fresh.decode_code(torch.zeros(1, h, h, dtype=torch.long), torch.zeros(1, 2 * h, 2 * h, dtype=torch.long))
    with pytest.raises(H.HQError):
        fresh.load_param("decoder.conv_in.weight", torch.zeros(3, 3))


def test_image_gpt2_sample_returns_pixels_with_stage1():
    """`ImageGPT2.sample` end to end: sampler + stage-1 decoder built from one config (random init), pixels in [0, 1]."""
    import os
    import hqtransformer_b200 as H
    path = os.path.join(os.path.dirname(H.__file__), "configs", "imagenet_l12.yaml")
    model = H.ImageGPT2.from_config(path, with_stage1=True, stage1_max_batch=4, device=0, precision="bf16", max_batch=4).eval()
    px = model.sample(cls_idx=7, top_k=256, num_candidates=4, is_tqdm=False)
    assert tuple(px.shape) == (4, 3, 256, 256) and float(px.min()) >= 0.0 and float(px.max()) <= 1.0
