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


@pytest.mark.parametrize("B", [8, 7, 1])          # B = 1: rank 1's shard is empty and it still joins the collective
def test_gather_codes_two_ranks_gloo(tmp_path, B):
    world, S = 2, 16
    port = 29600 + (os.getpid() + B) % 300
    mp.spawn(_worker, args=(world, port, B, S, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(torch.load(os.path.join(str(tmp_path), f"ok{r}.pt"))[0]) == 1


class _FakeModel:
    """Host-side stand-in: `sampling_ihqgpt_sharded` only needs `.device`; the sampler itself is patched out (no GPU)."""
    device = torch.device("cpu")


def _sharded_worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hqtransformer_b200.sampling as S
    from hqtransformer_b200.distributed import sampling_ihqgpt_sharded
    seen = {}

    def fake(model, n, cond, row_offset=0, seed=None, max_seq_len=256, **kw):
        seen["seed"], seen["n"], seen["lo"] = seed, n, row_offset
        rows = torch.arange(row_offset, row_offset + n, dtype=torch.int64)
        return (rows[:, None] * 1000 + seed % 7).expand(n, max_seq_len).contiguous(), \
            (rows[:, None, None] * 10).expand(n, max_seq_len, 4).contiguous()
    S.sampling_ihqgpt = fake
    torch.manual_seed(100 + rank)                      # ranks whose generators disagree must still share ONE key
    ct, cb = sampling_ihqgpt_sharded(_FakeModel(), B, 3, max_seq_len=4)
    want = torch.arange(B, dtype=torch.int64)[:, None] * 1000 + seen["seed"] % 7 if "seed" in seen else None
    ok = tuple(ct.shape) == (B, 4) and tuple(cb.shape) == (B, 4, 4)
    seeds = [None] * world
    dist.all_gather_object(seeds, seen.get("seed"))
    got = [x for x in seeds if x is not None]
    ok = ok and len(set(got)) == 1 and torch.equal(ct, (torch.arange(B)[:, None] * 1000 + got[0] % 7).expand(B, 4))
    torch.save(torch.tensor([int(ok)]), os.path.join(out_dir, f"ok{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 1])
def test_sharded_entry_shares_one_seed_and_survives_empty_shards(tmp_path, B):
    """world_size 2 on gloo: the default Philox key is drawn once (rank 0) and broadcast; with B = 1 rank 1 has an empty
    shard, skips the run and still joins the all-gather (no hang)."""
    world = 2
    port = 29900 + (os.getpid() + B) % 90
    mp.spawn(_sharded_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(torch.load(os.path.join(str(tmp_path), f"ok{r}.pt"))[0]) == 1
