"""World-size-2 gloo test (CPU) of the N>1 path: a batch of independent registrations is split by index across the
ranks (no data-path collective), results are gathered, timings reduced with MAX.  The per-rank compute is the CPU oracle
here (the GPU engine needs a device); the sharding / gathering code is the one bench.py and the batched API use."""
import os

import numpy as np
import torch.multiprocessing as mp

N_PAIRS = 5


def _pairs():
    from open3d_slam_b200 import synth
    rng = np.random.default_rng(0)
    src, tgt, nrm, _ = synth.planar_cloud_config1(n=400, noise=0.01)
    out = []
    for _k in range(N_PAIRS):
        d = synth.se3(*rng.uniform(-0.02, 0.02, 3), rng.uniform(-0.05, 0.05, 3))
        out.append((src @ d[:3, :3].T + d[:3, 3], tgt, nrm))
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist_
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist_.init_process_group("gloo", rank=rank, world_size=world)
    from open3d_slam_b200 import dist
    from oracle import oracle as O
    pairs = _pairs()
    mine = dist.shard_range(len(pairs), world, rank)
    local = []
    for i in mine:
        s, t, n = pairs[i]
        r = O.registration_icp_p2plane(s, t, n, 0.6, np.eye(4), max_iter=30)
        local.append(np.r_[r.T.reshape(-1), r.fitness, r.inlier_rmse, r.n_corr])
    table = dist.gather_results(np.array(local).reshape(len(local), 19), len(pairs), world, rank)
    tmax = dist.max_over_ranks(10.0 + rank, world)
    dist_.barrier()
    q.put((rank, table, tmax, list(mine)))
    dist_.destroy_process_group()


def test_two_ranks_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import oracle as O
    ref = []
    for s, t, n in _pairs():
        r = O.registration_icp_p2plane(s, t, n, 0.6, np.eye(4), max_iter=30)
        ref.append(np.r_[r.T.reshape(-1), r.fitness, r.inlier_rmse, r.n_corr])
    ref = np.array(ref)
    shards = {}
    for rank, table, tmax, mine in got:
        assert table.shape == (N_PAIRS, 19) and np.abs(table - ref).max() < 1e-12     # every rank sees the full, ordered table
        assert tmax == 11.0                                                           # MAX over ranks
        shards[rank] = mine
    assert sorted(shards[0] + shards[1]) == list(range(N_PAIRS)) and not set(shards[0]) & set(shards[1])


# ----------------------------------------------------------------------------------------------------------------------
# shared registration targets: built once by the owner, broadcast to the ranks that register against them
# ----------------------------------------------------------------------------------------------------------------------
N_UNITS = 5


def _unit(u):
    rng = np.random.default_rng(100 + u)
    n = 50 + 17 * u
    return rng.normal(size=(n, 3)), rng.normal(size=(n, 3))


def _bcast_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist_
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist_.init_process_group("gloo", rank=rank, world_size=world)
    from open3d_slam_b200 import dist
    local = {u: tuple(torch.from_numpy(a) for a in _unit(u)) for u in range(N_UNITS) if dist.owner_of(u, world) == rank}
    needed = [0, 1, 4] if rank == 0 else [1, 2, 3]
    got, recv = dist.broadcast_point_sets(local, N_UNITS, world, rank, None, needed)
    dist_.barrier()
    q.put((rank, {u: (x.numpy().copy(), n.numpy().copy()) for u, (x, n) in got.items()}, recv, sorted(local)))
    dist_.destroy_process_group()


def test_two_ranks_broadcast_shared_targets():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, sets, recv, owned in got:
        assert owned == [u for u in range(N_UNITS) if u % 2 == rank]                   # round-robin ownership
        assert sorted(sets) == ([0, 1, 4] if rank == 0 else [1, 2, 3])                 # only what the rank registers against
        for u, (x, n) in sets.items():
            rx, rn = _unit(u)
            assert np.array_equal(x, rx) and np.array_equal(n, rn)                      # bit-exact payload
        foreign = [u for u in range(N_UNITS) if u % 2 != rank]
        assert recv == sum(_unit(u)[0].size * 8 * 2 for u in foreign)                  # bytes that crossed the wire into this rank
