"""The N > 1 path on the CPU: gloo ranks run the PRODUCT's sharding front
(chomp_b200.distributed.ShardedEngine: shard, pad uneven shards, one all-gather,
double-buffered pipeline) with a deterministic per-point function standing in for
the GPU stages; the gathered table must equal the unsharded one bit for bit
(SURVEY.md section 8(e)).  The same front with the real stages runs in
tests/test_gpu_distributed.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chomp_b200 import design, distributed


def _stand_in(cosmo, halo, hod, n_theta=30):
    """No cross-point dependence, like the real path; status flags vary per point."""
    cosmo, halo, hod = (np.asarray(a) for a in (cosmo, halo, hod))
    base = cosmo[:, :1]*3.0 + halo[:, 2:3]*0.1 + hod[:, :1]*0.01
    table = torch.as_tensor(base + np.arange(n_theta)[None, :]*1e-3)
    status = torch.as_tensor((np.floor(hod[:, 1]*1000.0) % 3).astype(np.int32))
    return table, status


def _worker(rank, world, port, n_points, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = distributed.ShardedEngine(evaluate=_stand_in)
    assert (sharded.rank, sharded.world) == (rank, world)
    cosmo, halo, hod = design.synthetic_batch(n_points)
    w, st = sharded.wtheta(cosmo, halo, hod)
    # three pipelined steps over different batches
    batches = [design.synthetic_batch(n_points, seed=100 + s) for s in range(3)]
    piped = [(a.numpy().copy(), b.numpy().copy()) for a, b in sharded.pipeline(batches)]
    dist.barrier()
    if rank == 0:
        np.savez(out_path, w=w.numpy(), st=st.numpy(), **{"p%d" % i: p[0] for i, p in enumerate(piped)},
                 **{"s%d" % i: p[1] for i, p in enumerate(piped)})
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_points", [(2, 64), (2, 67), (3, 100)])
def test_sharded_all_gather_equals_single_rank(tmp_path, world, n_points):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "w.npz")
    mp.spawn(_worker, args=(world, port, n_points, out), nprocs=world, join=True)
    got = np.load(out)
    cosmo, halo, hod = design.synthetic_batch(n_points)
    table, status = _stand_in(cosmo, halo, hod)
    assert np.array_equal(got["w"], table.numpy())
    assert np.array_equal(got["st"], status.numpy())
    for s in range(3):
        t2, s2 = _stand_in(*design.synthetic_batch(n_points, seed=100 + s))
        assert np.array_equal(got["p%d" % s], t2.numpy())
        assert np.array_equal(got["s%d" % s], s2.numpy())


def test_shard_bounds_cover_the_batch():
    for n, world in ((64, 8), (67, 8), (5, 8), (4096, 3)):
        b = distributed.shard_bounds(n, world)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        sizes = [y - x for x, y in b]
        assert max(sizes) - min(sizes) <= 1


def test_single_process_front_needs_no_process_group():
    sharded = distributed.ShardedEngine(evaluate=_stand_in)
    cosmo, halo, hod = design.synthetic_batch(9)
    w, st = sharded.wtheta(cosmo, halo, hod)
    t, s = _stand_in(cosmo, halo, hod)
    assert np.array_equal(w.numpy(), t.numpy()) and np.array_equal(st.numpy(), s.numpy())
