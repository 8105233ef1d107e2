"""not gpu: the N>1 host path (batch sharding + the final all-gather of code grids) with 2 gloo ranks on CPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, S, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hqtransformer_b200.distributed import gather_codes, shard_range
    full_t = torch.arange(B * S, dtype=torch.int64).view(B, S) * 3 + 1
    full_b = torch.arange(B * S * 4, dtype=torch.int64).view(B, S, 4) * 7 + 2
    lo, hi = shard_range(B, rank, world)
    ct, cb = gather_codes(full_t[lo:hi].clone(), full_b[lo:hi].clone(), B)
    ok = torch.equal(ct, full_t) and torch.equal(cb, full_b)
    torch.save(torch.tensor([int(ok)]), os.path.join(out_dir, f"ok{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_gather_codes_two_ranks_gloo(tmp_path, B):
    world, S = 2, 16
    port = 29600 + (os.getpid() + B) % 300
    mp.spawn(_worker, args=(world, port, B, S, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(torch.load(os.path.join(str(tmp_path), f"ok{r}.pt"))[0]) == 1
