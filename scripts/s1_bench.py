"""Stage-1 decode throughput: the shipped HQ-VAE decoder (256 x 256) on random grids, random-init weights.
Prints images/s, ms per image, achieved conv TFLOP/s (2 * MACs of the convolutions, interior pixels) vs the measured bf16 peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hqtransformer_b200 as H  # noqa: E402

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.isfile(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
for B in [int(v) for v in (sys.argv[1:] or ["1", "8", "32"])]:
    dec = H.HQVAEDecoder(max_batch=B)
    dec.init_weights(seed=1)
    ct = torch.randint(0, 8192, (B, 8, 8), device="cuda")
    cb = torch.randint(0, 8192, (B, 16, 16), device="cuda")
    for _ in range(3):
        px = dec.decode_code(ct, cb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10 if B <= 8 else 5
    e0.record()
    for _ in range(n):
        px = dec.decode_code(ct, cb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = dec.last_conv_flops / (ms * 1e-3) * 1e-12
    print(json.dumps({"batch": B, "ms_per_batch": round(ms, 3), "ms_per_image": round(ms / B, 4), "images_per_s": round(B / ms * 1e3, 1),
                      "conv_gflop_per_image": round(dec.last_conv_flops / B * 1e-9, 1), "conv_tflops": round(tf, 1),
                      "frac_of_measured_bf16_sustained": round(tf / peaks["bf16_tflops_sustained"], 3),
                      "device_mb": round(dec.device_bytes / 2**20)}), flush=True)
    dec.close()
    del dec
    torch.cuda.empty_cache()

if os.environ.get("S1_REFERENCE"):
    # the UNMODIFIED reference decoder (baseline/_ref copy through oracle/ref_shim.py) on the same GPU, PyTorch / cuDNN:
    # one image at a time as measure_throughput does (measure_throughput/__main__.py:108-111), and batched
    from oracle import ref_shim as R, s1_oracle as S1
    cfg = S1.IMAGENET_S1
    P = S1.make_params(cfg, seed=2)
    model = R.build_reference_stage1(cfg, P).cuda()
    for B, chunked in [(8, True), (8, False), (32, False)]:
        ct = torch.randint(0, 8192, (B, 8, 8), device="cuda")
        cb = torch.randint(0, 8192, (B, 16, 16), device="cuda")

        def run():
            with torch.no_grad():
                if chunked:
                    return torch.cat([model.decode_code(a, b) for a, b in zip(ct.chunk(B), cb.chunk(B))], 0)
                return model.decode_code(ct, cb)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"impl": "reference decode_code on the same GPU (PyTorch eager, fp32 / TF32 cuDNN)", "batch": B,
                          "one_image_at_a_time": chunked, "ms_per_image": round(ms / B, 4), "images_per_s": round(B / ms * 1e3, 1)}), flush=True)
