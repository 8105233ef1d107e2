"""Device timeline of the sampling loop (GPU box): per-kernel durations and gaps for a few top positions."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hqtransformer_b200 as H
from hqtransformer_b200.engine import SamplingParams

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P0, P1 = 30, 34
for a in sys.argv:
    if a.startswith("--pos="):
        P0 = int(a.split("=")[1]); P1 = P0 + 4
graph = "--no-graph" not in sys.argv
pdl = "--no-pdl" not in sys.argv
cfg = os.path.join(os.path.dirname(H.__file__), "configs", "imagenet_l12.yaml")
conf = H.load_config(cfg)
for a in sys.argv:
    if a.startswith("--layers="):
        conf.stage2.hparams.n_layers = int(a.split("=")[1])
model = H.ImageGPT2(conf, device=0, precision="bf16", max_batch=B, use_cuda_graph=graph, use_pdl=pdl)
model.stage2.init_weights(0)
s2 = model.stage2
eng = s2.engine("bf16")
cond = torch.randint(0, 1000, (B,), device="cuda")
ct, cb = H.sampling_ihqgpt(s2, B, cond, max_seq_len=64, is_tqdm=False)     # fills codes + cache
torch.cuda.synchronize()
for rep in range(2):   # second pass = warm
    tl = eng.trace_run(batch=B, seq_len=64, pos_begin=P0, pos_end=P1, sampling=SamplingParams(), cond=cond,
                       codes_top=ct, codes_bot=cb)
t0 = min(s for _, s, _ in tl)
n_pos = P1 - P0
agg = collections.OrderedDict()
prev_end = None
total_gap = 0
for tag, s, e in tl:
    d = agg.setdefault(tag, [0, 0.0, 0.0])
    d[0] += 1
    d[1] += (e - s)
    if prev_end is not None:
        gap = s - prev_end
        d[2] += gap
        total_gap += gap
    prev_end = max(prev_end or 0, e)
span = max(e for _, _, e in tl) - t0
print(f"B={B} graph={graph} pdl={pdl}: {len(tl)} launches over {n_pos} positions, span {span/1e3/n_pos:.1f} us/position, "
      f"sum of kernel lifetimes {sum(v[1] for v in agg.values())/1e3/n_pos:.1f} us/position, sum of gaps {total_gap/1e3/n_pos:.1f} us/position")
e2e = collections.OrderedDict()
prev = None
for tag, s, e in tl:
    if prev is not None:
        d = e2e.setdefault(tag, [0, 0.0])
        d[0] += 1
        d[1] += e - prev
    prev = e
print(f"{'kernel':34s} {'n/pos':>6s} {'avg_us':>8s} {'gap_before':>10s} {'us/pos':>8s} {'end-to-end delta us':>20s}")
for tag, (n, dur, gap) in agg.items():
    dd = e2e.get(tag, [1, 0.0])
    print(f"{tag:34s} {n/n_pos:6.1f} {dur/n/1e3:8.2f} {gap/n/1e3:10.2f} {dur/1e3/n_pos:8.1f} {dd[1]/dd[0]/1e3:20.2f}")
json.dump([(t, s - t0, e - t0) for t, s, e in tl], open(f"gpurun_out/timeline_B{B}_g{int(graph)}_p{int(pdl)}_L{conf.stage2.hparams.n_layers}.json", "w"))
