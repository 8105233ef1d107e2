"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : batch-sharded sampling over N GPUs (NCCL all-gather of the
code grids) must equal the single-GPU result of the whole batch, for greedy AND stochastic sampling."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import hqtransformer_b200 as H
from oracle import hq_oracle as O
from tests.helpers import build_model


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = O.SMALL
    P = O.make_params(cfg, seed=5, init="rich")
    import tests.helpers as TH
    B = 37                                   # ragged over the ranks on purpose
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, cfg.n_classes, (B,), generator=g)
    # every rank: model on its own GPU
    _orig = H.iHQGPT.__init__
    model = TH.build_model(cfg, P, precision="bf16", max_batch=B, max_seq_len=16) if local == 0 else None
    if local != 0:
        model = H.iHQGPT(vocab_size_top=cfg.vocab_top, vocab_size_bot=cfg.vocab_bot, vocab_size_txt=cfg.vocab_txt,
                         ratio_bot2top=4, use_cls_cond=True, use_txt_cond=False, model_type="parallel",
                         hparams=TH.hparams_of(cfg), hparams_dec=TH.hparams_of(cfg, cfg.n_layers_depth), device=local,
                         precision="bf16", max_batch=B, max_seq_len=16)
        model.load_state_dict(P, strict=True)
    ok = True
    for kw in (dict(top_k_top=1, top_k_bot=1), dict(top_k_top=50, top_p_top=0.9, top_k_bot=50, top_p_bot=0.9,
                                                     softmax_temperature=[0.9, 0.9])):
        ct, cb = H.sampling_ihqgpt_sharded(model, B, labels.cuda(), max_seq_len=16, is_tqdm=False, seed=3, **kw)
        ct1, cb1 = H.sampling_ihqgpt(model, B, labels.cuda(), max_seq_len=16, is_tqdm=False, seed=3, **kw)
        ok = ok and torch.equal(ct, ct1) and torch.equal(cb, cb1) and tuple(ct.shape) == (B, 16)
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if int(flag) == 1 else "MULTI_GPU_MISMATCH", f"world={world}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
