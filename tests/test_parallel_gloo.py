"""world_size-2 gloo test of the sharding + single gather (the N > 1 path of bench.py), CPU only."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from laplacesolvermps_b200.parallel import shard_range, gather_results
    lo, hi = shard_range(n_total, rank, world)
    local = torch.stack([torch.arange(lo, hi, dtype=torch.float64), torch.arange(lo, hi, dtype=torch.float64) ** 2], dim=1)
    full = gather_results(local, n_total)
    ok = full.shape == (n_total, 2) and torch.equal(full[:, 0], torch.arange(n_total, dtype=torch.float64)) \
        and torch.equal(full[:, 1], torch.arange(n_total, dtype=torch.float64) ** 2)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, 11, ret), nprocs=2, join=True)
        assert ret[0] and ret[1]
