#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
TMO=400 TAILN=30 run r2_f4_tests python -m pytest tests/test_gpu_sampling_loop.py -q -p no:cacheprovider -k "shared_text_prefix or text_prefix"
TMO=300 TAILN=6 run r2_f4_bench python - <<'PY'
import os, sys, time, json, torch
sys.path.insert(0, os.getcwd())
import hqtransformer_b200 as H
cfg = os.path.join("hqtransformer_b200", "configs", "cc15m_l12.yaml")
B = 256
m = H.ImageGPT2.from_config(cfg, device=0, precision="bf16", max_batch=B).eval()
ids = torch.randint(0, 16384, (1, 64), device="cuda").repeat(B, 1)
kw = dict(top_k_top=2048, top_k_bot=2048, max_seq_len=64, is_tqdm=False, use_fp16=True)
for shared in (False, True):
    for i in range(2):
        H.sampling_ihqgpt(m.stage2, B, ids, seed=i, shared_prefix=shared, **kw)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        H.sampling_ihqgpt(m.stage2, B, ids, seed=5 + i, shared_prefix=shared, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"model": "cc15m_l12", "batch": B, "one_prompt_for_all_rows": True, "shared_prefix": shared, "ms_per_step": round(ms, 2), "images_per_s": round(B / ms * 1e3, 1)}), flush=True)
PY
