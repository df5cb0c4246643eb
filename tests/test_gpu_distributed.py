"""The product's multi-GPU front (chomp_b200.distributed.ShardedEngine) with the real CUDA stages:
G = 1 and G = N must give bit-identical gathered tables.

* one GPU: the shards of G = 2, 3, 8 emulated ranks, evaluated one after the other on separate handles,
  concatenate to exactly the unsharded table (a point's result does not depend on its batch);
* two or more GPUs: two NCCL ranks, uneven split, gathered table compared bit for bit with the unsharded
  evaluation on rank 0 (skipped on a one-GPU box; the driver's scaling run covers N = 2, 4, 8)."""
import os
import socket

import numpy as np
import pytest

from chomp_b200 import design, distributed, engine

pytestmark = pytest.mark.gpu


def _survey():
    return engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1), bins_per_decade=10.0,
                         power_spec="power_gg")


def test_emulated_ranks_concatenate_to_the_unsharded_table():
    import torch
    survey = _survey()
    n = 67
    cosmo, halo, hod = design.synthetic_batch(n)
    one = distributed.ShardedEngine(survey)
    assert one.world == 1
    w_full, st_full = one.wtheta(cosmo, halo, hod)
    w_full, st_full = w_full.cpu().numpy(), st_full.cpu().numpy()
    assert w_full.shape == (n, 30) and np.all(np.isfinite(w_full)) and not st_full.any()
    w_host, st_host = one.wtheta_host(cosmo, halo, hod)
    assert np.array_equal(w_host, w_full) and np.array_equal(st_host, st_full)
    for world in (2, 3, 8):
        parts = []
        for rank, (a, b) in enumerate(distributed.shard_bounds(n, world)):
            rank_engine = distributed.ShardedEngine(survey)      # a handle of its own, like another process
            w, st = rank_engine._evaluate_engine(torch.as_tensor(cosmo[a:b]).cuda(), torch.as_tensor(halo[a:b]).cuda(),
                                                 torch.as_tensor(hod[a:b]).cuda())
            parts.append(w.cpu().numpy())
        assert np.array_equal(np.concatenate(parts, 0), w_full), world
    # the double-buffered pipeline returns the same tables, in order
    batches = [design.synthetic_batch(n, seed=7 + s) for s in range(3)]
    piped = [w.cpu().numpy() for w, st in one.pipeline(batches)]
    for w, batch in zip(piped, batches):
        ref, _ = one.wtheta(*batch)
        assert np.array_equal(w, ref.cpu().numpy())


def _nccl_worker(rank, world, port, n_points, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    survey = _survey()
    sharded = distributed.ShardedEngine(survey, device=rank)
    cosmo, halo, hod = design.synthetic_batch(n_points)
    w, st = sharded.wtheta(cosmo, halo, hod)
    w_host, st_host = sharded.wtheta_host(cosmo, halo, hod)
    ok = np.array_equal(w.cpu().numpy(), w_host) and np.array_equal(st.cpu().numpy(), st_host)
    # config 5 through the same front: w(theta) | covariance | status in ONE all-gather (uneven split of 7 points)
    from chomp_b200 import _lib
    cov_front = distributed.ShardedEngine(survey, device=rank, tri_moment=_lib.TRISPECTRUM_MOMENT["power_gggg"])
    setup = engine.CovarianceSetup(survey, (0.001, 1.0), 10.0, 25.0, [1e10, 1e10], [1e10, 1e10], 1.0, True, "power_gg")
    wc, cov, stc = cov_front.covariance(cosmo[:7], halo[:7], hod[:7], setup)
    wc_h, cov_h, stc_h = cov_front.covariance_host(cosmo[:7], halo[:7], hod[:7], setup)
    ok = ok and np.array_equal(cov.cpu().numpy(), cov_h) and np.array_equal(wc.cpu().numpy(), wc_h) and not stc_h.any()
    if rank == 0:
        single, _ = sharded._evaluate_engine(torch.as_tensor(cosmo).cuda(), torch.as_tensor(halo).cuda(),
                                             torch.as_tensor(hod).cuda())
        cov_single = cov_front.engine.covariance(cosmo[:7], halo[:7], hod[:7], setup)
        np.savez(out_path, w=w.cpu().numpy(), single=single.cpu().numpy(), ok=ok, cov=cov.cpu().numpy(),
                 cov_single=cov_single.cpu().numpy(), wc=wc.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_nccl_ranks_equal_one_rank(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the one-GPU variant above emulates the ranks)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "w.npz")
    mp.spawn(_nccl_worker, args=(2, port, 67, out), nprocs=2, join=True)
    got = np.load(out)
    assert bool(got["ok"])
    assert np.array_equal(got["w"], got["single"])
    assert np.array_equal(got["cov"], got["cov_single"]) and got["cov"].shape == (7, 30, 30)
    assert np.array_equal(got["wc"], got["single"][:7])
