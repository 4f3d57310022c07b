"""N > 1 host logic on CPU: two gloo ranks shard the points, each computes the statistics of its slice (with the oracle --
the GPU is not available here), the packed buffers are sum-all-reduced exactly as libsgp does with NCCL, and every rank must
end up with the statistics of the full data set."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussianprocessnode_b200 import shard
from oracle import batched


def test_shard_bounds_partition():
    for N in (0, 1, 31, 32, 1000, 10_000_000):
        for world in (1, 2, 3, 8):
            b = [shard.shard_bounds(N, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_bounds(10, 2, 2)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    M = 7
    A = rng.normal(size=(M, M)); psi2 = A @ A.T
    psi1 = rng.normal(size=M)
    buf = shard.pack_stats(1.5, psi1, psi2, 2.5, 3.5, 42)
    p0, p1, p2, sy, sw, n = shard.unpack_stats(buf, M)
    assert p0 == 1.5 and sy == 2.5 and sw == 3.5 and n == 42
    assert np.array_equal(p1, psi1) and np.array_equal(p2, psi2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    N, D, M = 1001, 3, 24                      # ragged: 501 + 500 points
    X = rng.normal(size=(N, D)); y = rng.normal(size=N); yv = rng.uniform(0.0, 0.2, N); w = rng.uniform(0.5, 1.5, N)
    Z = X[:M].copy(); ell = np.array([1.0, 2.0, 0.7])
    lo, hi = shard.shard_bounds(N, world, rank)
    p0, p1, p2, sy = batched.psi_stats_point(X[lo:hi], y[lo:hi], Z, 1.3, ell, yvar=yv[lo:hi], weights=w[lo:hi])
    t = torch.from_numpy(shard.pack_stats(p0, p1, p2, sy, w[lo:hi].sum(), hi - lo))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)       # the one exchange of a sharded sweep
    q0, q1, q2, qy, qw, qn = shard.unpack_stats(t.numpy(), M)
    f0, f1, f2, fy = batched.psi_stats_point(X, y, Z, 1.3, ell, yvar=yv, weights=w)
    ok = (qn == N and abs(q0 - f0) <= 1e-12 * abs(f0) and abs(qy - fy) <= 1e-12 * abs(fy) and abs(qw - w.sum()) <= 1e-12 * w.sum()
          and np.linalg.norm(q2 - f2) <= 1e-13 * np.linalg.norm(f2) and np.linalg.norm(q1 - f1) <= 1e-13 * np.linalg.norm(f1))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_full_sweep():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_packed_lower_roundtrip_and_shares():
    rng = np.random.default_rng(1)
    for M, D_out in ((1, 1), (7, 1), (8, 3), (33, 2)):
        A = rng.normal(size=(M, M)); psi2 = A @ A.T; psi1 = rng.normal(size=(M, D_out))
        buf = shard.pack_stats_lower(1.5, psi1, psi2, 2.5, 3.5, 42)
        assert buf.size % 2 == 0 and buf.size >= M * (M + 1) // 2 + M * D_out + 4
        p0, p1, p2, sy, sw, n = shard.unpack_stats_lower(buf, M, D_out)
        assert (p0, sy, sw, n) == (1.5, 2.5, 3.5, 42) and np.array_equal(p2, psi2) and np.array_equal(np.asarray(p1).reshape(M, D_out), psi1)
        for world in (1, 2, 3, 8):
            sh = [shard.two_shot_share(buf.size, world, r) for r in range(world)]
            assert sh[0][0] == 0 and sh[-1][1] == buf.size and all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))


def _worker_two_shot(rank, world, port, out):
    """The exchange of csrc/xchg.cuh restated on CPU tensors: contributions in the packed layout, reduce-scatter of every rank's share in rank
    order, shares delivered to everybody, expansion -- bitwise identical results on all ranks, equal to the full sweep."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    N, D, M = 777, 2, 19
    X = rng.normal(size=(N, D)); y = rng.normal(size=N)
    Z = X[:M].copy(); ell = np.array([1.1, 0.8])
    lo, hi = shard.shard_bounds(N, world, rank)
    p0, p1, p2, sy = batched.psi_stats_point(X[lo:hi], y[lo:hi], Z, 0.9, ell)
    mine = torch.from_numpy(shard.pack_stats_lower(p0, p1, p2, sy, float(hi - lo), hi - lo))
    contrib = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(contrib, mine)                               # "every rank can read every contribution buffer"
    a, b = shard.two_shot_share(mine.numel(), world, rank)
    share = contrib[0][a:b].clone()
    for q in range(1, world):
        share += contrib[q][a:b]                                 # rank order: every element is summed exactly once
    sizes = [shard.two_shot_share(mine.numel(), world, q) for q in range(world)]
    parts = [torch.empty(hi_ - lo_, dtype=torch.float64) for lo_, hi_ in sizes]
    dist.all_gather(parts, share) if len({p.numel() for p in parts}) == 1 else [dist.broadcast(parts[q] if q != rank else share, src=q) for q in range(world)]
    parts[rank] = share
    res = torch.cat(parts).numpy()
    q0, q1, q2, qy, qw, qn = shard.unpack_stats_lower(res, M)
    f0, f1, f2, fy = batched.psi_stats_point(X, y, Z, 0.9, ell)
    ok = (qn == N and np.linalg.norm(q2 - f2) <= 1e-13 * np.linalg.norm(f2) and np.linalg.norm(q1 - f1) <= 1e-13 * np.linalg.norm(f1)
          and abs(q0 - f0) <= 1e-12 * abs(f0) and abs(qy - fy) <= 1e-12 * abs(fy) and np.array_equal(q2, q2.T))
    import hashlib
    out[rank] = (bool(ok), hashlib.sha256(res.tobytes()).hexdigest())
    dist.destroy_process_group()


def test_two_rank_gloo_two_shot_packed_exchange():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_two_shot, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] and res[1][0]
    assert res[0][1] == res[1][1], "every rank must hold bitwise identical statistics"
