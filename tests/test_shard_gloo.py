"""world_size-2 `gloo` test of the N>1 host logic: contiguous sharding of a batch over
ranks (no data-path collective) and the max-over-ranks timing reduction bench.py uses."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_items, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import _b200pkg
    pkg = _b200pkg.load()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lo, hi = pkg.shard_range(n_items, rank, world)
    owned = torch.zeros(n_items, dtype=torch.int64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)                       # every item owned exactly once
    t = torch.tensor([float(10 + rank)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # bench.py: time = max over ranks
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([lo, hi, int(owned.min()), int(owned.max()), t.item()]))
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 10001, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert r0[0] == 0 and r0[1] == r1[0] and r1[1] == 10001
    assert r0[2] == r0[3] == 1
    assert r0[4] == r1[4] == 11.0
