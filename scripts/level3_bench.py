"""3-level HQTransformer sampling throughput on one B200: ImageNet level-3 architecture (D = 1536, 12 + 4
layers, 3 x 8192 codes; 8x8 top + 16x16 middle + 32x32 bottom codes per image), random init, measure_throughput protocol
(top-k / top-p None, T = 1), plus the unmodified reference sampler (fp16 autocast, PyTorch eager) on the same GPU."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hqtransformer_b200 as H  # noqa: E402

cfg_path = os.path.join(ROOT, "hqtransformer_b200", "configs", "imagenet_l12_level3.yaml")
for B in [int(v) for v in (sys.argv[1:] or ["64", "256"])]:
    model = H.ImageGPT2.from_config(cfg_path, device=0, precision="bf16", max_batch=B).eval()
    s2 = model.stage2
    cond = torch.randint(0, 1000, (B,), device="cuda")
    kw = dict(top_k=[None] * 3, top_p=[None] * 3, softmax_temperature=[1.0] * 3, use_fp16=True, is_tqdm=False, max_seq_len=64)
    for i in range(2):
        H.sampling_hqtransformer(s2, B, cond, seed=i, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for i in range(n):
        codes = H.sampling_hqtransformer(s2, B, cond, seed=10 + i, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"model": "imagenet_l12_level3", "batch": B, "images_per_s": round(B / ms * 1e3, 1), "ms_per_step": round(ms, 2),
                      "ms_per_top_position": round(ms / 64, 3), "codes_per_image": 64 * 21,
                      "launches_per_run": s2.engine("bf16").last_launch_count}), flush=True)
    del model, s2
    torch.cuda.empty_cache()

if os.environ.get("L3_REFERENCE"):
    from oracle import hq3_oracle as O3, ref_shim as R
    if R.reference_available():
        import contextlib
        import io
        R.import_reference()
        from hqvae.models.stage2.hqtransformer import HQTransformer as RefHQ
        cfg = O3.IMAGENET_L12_LEVEL3
        hp = R.make_hparams(cfg.embed_dim, cfg.n_layers, cfg.n_heads, n_classes=1000, ctx_len_img=256)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = RefHQ(vocab_sizes=[8192] * 3, vocab_size_txt=16384, decoding_type="parallel-add", use_cls_cond=True,
                        use_txt_cond=False, hparams=hp, hparams_dec=None).cuda().eval()
        _, S = R.import_reference()
        B = int(os.environ.get("L3_REFERENCE_B", 64))

        def one():
            return S.sampling_hqtransformer(ref, num_candidates=B, cond=7, top_k=[None] * 3, top_p=[None] * 3,
                                            softmax_temperature=[1.0] * 3, is_tqdm=False, use_fp16=True, max_seq_len=64)
        with torch.no_grad():
            one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"impl": "unmodified reference HQTransformer sampler (PyTorch eager, fp16 autocast) on the same GPU",
                          "batch": B, "images_per_s": round(B / ms * 1e3, 1), "ms_per_top_position": round(ms / 64, 3)}), flush=True)
