"""The N > 1 path on the CPU: two gloo ranks shard one parameter batch, each
"evaluates" its slice, one all-gather assembles the result tables; the gathered
table must equal the unsharded one bit for bit."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chomp_b200 import design


def _fake_wtheta(cosmo, halo, hod, n_theta=30):
    """A deterministic per-point function standing in for the GPU path (no
    cross-point dependence, like the real one)."""
    base = cosmo[:, :1]*3.0 + halo[:, 2:3]*0.1 + hod[:, :1]*0.01
    return torch.as_tensor(base + np.arange(n_theta)[None, :]*1e-3)


def _worker(rank, world, port, n_points, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cosmo, halo, hod = design.synthetic_batch(n_points)
    sl = design.shard(n_points, rank, world)
    local = _fake_wtheta(cosmo[sl], halo[sl], hod[sl])
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    full = torch.cat(gathered, 0)
    dist.barrier()
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.destroy_process_group()


def test_two_rank_all_gather_equals_single_rank(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_points = 64                       # divisible: all_gather needs equal shards
    out = str(tmp_path / "w.npy")
    mp.spawn(_worker, args=(2, port, n_points, out), nprocs=2, join=True)
    cosmo, halo, hod = design.synthetic_batch(n_points)
    expect = _fake_wtheta(cosmo, halo, hod).numpy()
    assert np.array_equal(np.load(out), expect)
