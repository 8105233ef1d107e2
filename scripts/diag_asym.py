"""Diagnostic: fp32 greedy parity of the ASYM golden under graph + PDL, repeated, with mismatch positions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hqtransformer_b200 as H
from oracle import hq_oracle as O
from tests.helpers import build_model, cfg_from_meta, load_golden
GREEDY = dict(top_k_top=1, top_p_top=1.0, top_k_bot=1, top_p_bot=1.0, softmax_temperature=[1.0, 1.0])
g, meta = load_golden("asym_cls_greedy.npz")
cfg = cfg_from_meta(meta)
P = O.make_params(cfg, seed=meta["seed"], init=meta["init"])
labels = torch.from_numpy(g["labels"])
def report(tag, ct, cb, wt, wb):
    ct, cb = ct.cpu().numpy(), cb.cpu().numpy()
    bad_t = np.argwhere(ct != wt); bad_b = np.argwhere(cb != wb)
    first = (bad_t[0].tolist() if len(bad_t) else None, bad_b[0].tolist() if len(bad_b) else None)
    print(f"{tag}: top mismatches {len(bad_t)} bottom mismatches {len(bad_b)} first {first}", flush=True)
for graph, pdl in [(True, True), (True, False), (False, True)]:
    for rep in range(3):
        model = build_model(cfg, P, precision="fp32", use_cuda_graph=graph, use_pdl=pdl)
        order = ["rows", "scalar", "rows", "scalar"] if rep != 1 else ["scalar", "rows", "scalar"]
        for what in order:
            if what == "rows":
                ct, cb = H.sampling_ihqgpt(model, len(labels), labels, use_fp16=False, max_seq_len=64, is_tqdm=False, **GREEDY)
                report(f"graph={graph} pdl={pdl} rep={rep} per-row labels", ct, cb, g["codes_top"], g["codes_bot"])
            else:
                ct, cb = H.sampling_ihqgpt(model, len(labels), int(labels[-1]), use_fp16=False, max_seq_len=64, is_tqdm=False, **GREEDY)
                report(f"graph={graph} pdl={pdl} rep={rep} scalar class  ", ct, cb, g["codes_top_scalar_class"], g["codes_bot_scalar_class"])
