"""World-size-2 gloo run of the multi-GPU host logic (instance sharding + final gather)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adacharge_b200 import sharding


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_items = 11
    mine = sharding.shard_range(n_items, rank, world)
    # stand-in for "solve my instances": a per-instance scalar
    local = torch.tensor([float(len(mine)), float(sum(i * i for i in mine)), float(max(mine)), float(rank)], dtype=torch.float64)
    gathered = sharding.gather_summaries(local)
    tot = torch.stack(gathered).numpy()
    slow = sharding.max_over_ranks(10.0 + rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([tot.ravel(), [slow]]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            parts = [sharding.shard_range(n, r, w) for r in range(w)]
            assert sum(len(p) for p in parts) == n
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gather(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = (np.load(tmp_path / f"r{r}.npy") for r in (0, 1))
    np.testing.assert_array_equal(a, b)  # every rank sees the same gathered summary
    tot = a[:-1].reshape(2, 4)
    assert tot[:, 0].sum() == 11 and tot[:, 1].sum() == sum(i * i for i in range(11))
    assert list(tot[:, 3]) == [0.0, 1.0] and a[-1] == 11.0  # max over ranks


def test_single_process_passthrough():
    t = torch.tensor([1.0, 2.0])
    assert sharding.gather_summaries(t)[0] is t
    assert sharding.max_over_ranks(3.5) == 3.5
