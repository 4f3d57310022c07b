"""Sharded sweep over 2 (and 4) GPUs through the C ABI: N is cut into slices, one process / sgp_ctx per GPU, the packed
statistics are summed over the ranks inside sgp_sweep_psi -- by the sweep kernel itself through NVLink peer memory (CUDA IPC,
two-shot all-reduce of the packed lower triangle in the kernel's tail; a kernel of its own for the sweeps without that tail), or by one
NCCL all-reduce when SGP_COMM_P2P=0; every rank must hold the single-GPU result (tolerance: the summation order differs, relative
Frobenius <= 1e-12) and all ranks the same bits.  Skipped on boxes with fewer GPUs (bench.py --gpus N carries the same check in
its `parity` block, which the driver's scaling run records)."""
import hashlib
import os
import socket
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out, p2p, M):
    os.environ["SGP_COMM_P2P"] = "1" if p2p else "0"
    import torch
    import torch.distributed as dist
    from gaussianprocessnode_b200 import SGPContext, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # host plumbing only: carries the NCCL unique id
    uid = [SGPContext.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    rng = np.random.default_rng(11)
    N, D = 50_001, 8                                              # M > 384: the generate-once kernel, which carries the fused exchange
    X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=N); yv = rng.uniform(0, 0.1, N)
    Z = X[:M].copy(); ell = np.full(D, 2.0)
    ctx = SGPContext(rank); ctx.set_kernel(1.2, ell); ctx.set_inducing(Z)
    sw = shard.ShardedSweep(ctx, world, rank, uid[0])
    sw.set_data(X, y, yv)
    p0, p1, p2, sy = sw.sweep_psi()
    q0, q1, q2, qy = sw.sweep_psi()                                  # a second sweep: the exchange epoch advances, same bits
    same = bool(np.array_equal(p2, q2) and np.array_equal(p1, q1) and p0 == q0 and sy == qy)
    digest = hashlib.sha256(p2.tobytes() + p1.tobytes()).hexdigest()
    info = ctx.last_sweep_info()
    # the one-call host step under sharding (collective): upload beside the launch, device-side ready word, then the exchange; Psi2 packed =
    # this rank's result buffer of the exchange (peer-memory path) or a packed copy of the all-reduced square (NCCL path, first fused kernel)
    from gaussianprocessnode_b200.sgp import pack_lower
    lo, hi = shard.shard_bounds(N, world, rank)
    for it in range(3):
        h0, h1, hk, hy = ctx.sweep_psi_host(X[lo:hi], y[lo:hi], yv[lo:hi], packed=True)
        g0, g1, g2, gy = ctx.sweep_psi_host(X[lo:hi], y[lo:hi], yv[lo:hi])
        same = same and bool(np.array_equal(hk, pack_lower(p2)) and np.array_equal(g2, p2) and np.array_equal(h1, p1) and np.array_equal(g1, p1)
                             and h0 == p0 == g0 and hy == sy == gy and np.array_equal(ctx.fetch_psi2_packed(), hk) and hk.base[-1] == N)
    ctx.close()
    single = SGPContext(rank); single.set_kernel(1.2, ell); single.set_inducing(Z); single.set_data(X, y, yv)
    f0, f1, f2, fy = single.sweep_psi(); single.close()
    out[rank] = (float(np.linalg.norm(p2 - f2) / np.linalg.norm(f2)), float(np.linalg.norm(p1 - f1) / np.linalg.norm(f1)),
                 float(abs(p0 - f0) / abs(f0)), float(abs(sy - fy) / abs(fy)), same, digest, info["launches"])
    dist.destroy_process_group()


@pytest.mark.parametrize("p2p,M", [(True, 600), (False, 600), (True, 300)])
def test_two_gpu_sharded_sweep_matches_single_gpu(p2p, M):
    # M = 300 goes to the first fused kernel, which has no exchange of its own: NCCL all-reduce on the (peer-mapped) statistics buffer
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out, p2p, M), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for r in res.values():
        assert max(r[:4]) <= 1e-12, res
        assert r[4], "repeated sweeps must give the same bits"
        assert r[6] == (1 if (p2p and M > 384) else 2), r              # fused exchange: ONE launch per sweep; NCCL path: kernel + all-reduce
    assert res[0][5] == res[1][5], "every rank must hold bitwise identical statistics"


def _worker_uncertain(rank, world, port, out, p2p):
    """MultiSGP statistics at the pendulum shape (GPnode/MultiSGPnode.jl:290-328 summed over the nodes of all ranks) and the theta step."""
    os.environ["SGP_COMM_P2P"] = "1" if p2p else "0"
    import torch
    import torch.distributed as dist
    from gaussianprocessnode_b200 import SGPContext, SRCUBATURE, GAUSSHERMITE
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = [SGPContext.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    rng = np.random.default_rng(124)
    N, d, M = 300, 2, 48
    means = rng.normal(size=(N, d)); A = rng.normal(size=(N, d, d)) * 0.1
    covs = A @ np.swapaxes(A, 1, 2) + 1e-2 * np.eye(d); R = rng.normal(size=(N, 2))
    gx, gy = np.meshgrid(np.linspace(-2.5, 2.5, 8), np.linspace(-2.5, 2.5, 6)); Z = np.stack([gx.ravel(), gy.ravel()], 1)
    ell = np.array([1.0, 1.2])
    lo, hi = N * rank // world, N * (rank + 1) // world
    ctx = SGPContext(rank); ctx.set_kernel(1.1, ell); ctx.set_inducing(Z); ctx.comm_init(world, rank, uid[0])
    single = SGPContext(rank); single.set_kernel(1.1, ell); single.set_inducing(Z)
    errs = []; digests = []
    for method, kw in ((SRCUBATURE, {}), (GAUSSHERMITE, {"p": 5})):
        for Dout, Rm in ((2, R), (1, R[:, :1])):
            a0, a1, a2, _ = ctx.sweep_psi_uncertain(method, means[lo:hi], covs[lo:hi], R=Rm[lo:hi], D_out=Dout, **kw)
            b0, b1, b2, _ = single.sweep_psi_uncertain(method, means, covs, R=Rm, D_out=Dout, **kw)
            errs += [float(np.linalg.norm(a2 - b2) / np.linalg.norm(b2)), float(np.linalg.norm(a1 - b1) / np.linalg.norm(b1)), float(abs(a0 - b0) / abs(b0))]
            digests.append(hashlib.sha256(a2.tobytes() + np.asarray(a1).tobytes()).hexdigest())
    # theta step: statistics summed over the ranks, rank-local gradient part summed at the end
    rng2 = np.random.default_rng(3)
    Nt, D8, Mt = 4001, 8, 96
    X = rng2.normal(size=(Nt, D8)); y = np.sin(X[:, 0]); Zt = X[:Mt].copy(); ell8 = np.full(D8, 1.8)
    v = rng2.normal(size=Mt) * 0.1; C = rng2.normal(size=(Mt, Mt)) * 0.05; Uv = np.linalg.cholesky(C @ C.T + 0.1 * np.eye(Mt)).T
    lo, hi = Nt * rank // world, Nt * (rank + 1) // world
    ctx.set_kernel(0.8, ell8); ctx.set_inducing(Zt); ctx.set_data(X[lo:hi], y[lo:hi])
    single.set_kernel(0.8, ell8); single.set_inducing(Zt); single.set_data(X, y)
    f = ctx.theta_objective(v, Uv, 30.0, 1e-6); g = single.theta_objective(v, Uv, 30.0, 1e-6)
    errs += [abs(f[0] - g[0]) / abs(g[0]), abs(f[1] - g[1]) / abs(g[1]), float(np.linalg.norm(f[2] - g[2]) / np.linalg.norm(g[2]))]
    # ... and the statistics left resident are the global ones: the :w terms agree with the single-GPU ones
    ctx.kuu_factor(1e-6, fetch=False); single.kuu_factor(1e-6, fetch=False)
    wa = ctx.w_terms(v, Uv); wb = single.w_terms(v, Uv)
    errs += [abs(wa[0] - wb[0]) / abs(wb[0]), abs(wa[1] - wb[1]) / abs(wb[1])]
    ctx.close(); single.close()
    out[rank] = (max(errs), digests)
    dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [True, False])
def test_two_gpu_uncertain_sweep_and_theta_step(p2p):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_uncertain, args=(2, _free_port(), out, p2p), nprocs=2, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    assert max(r[0] for r in res.values()) <= 1e-11, res
    assert res[0][1] == res[1][1], "every rank must hold bitwise identical statistics"


def test_four_gpu_sharded_sweep_matches_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 4:
        pytest.skip("needs four GPUs")
    world = 4
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out, True, 600), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == set(range(world))
    for r in res.values():
        assert max(r[:4]) <= 1e-12 and r[4] and r[6] == 1, res
    assert len({r[5] for r in res.values()}) == 1, "every rank must hold bitwise identical statistics"


def _worker_uneven(rank, world, port, out):
    """Ragged and EMPTY shards: 33 points over two ranks, then a single point (rank 1 owns nothing and still takes part in the sum)."""
    import torch
    import torch.distributed as dist
    from gaussianprocessnode_b200 import SGPContext, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = [SGPContext.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    rng = np.random.default_rng(2)
    errs = []
    ctx = SGPContext(rank); single = SGPContext(rank)
    first = True
    for M in (600, 40):
        for N in (33, 1):
            D = 3
            X = rng.normal(size=(N, D)); y = rng.normal(size=N); Z = rng.normal(size=(M, D)); ell = np.full(D, 1.3)
            ctx.set_kernel(1.0, ell); ctx.set_inducing(Z)
            if first:
                sw = shard.ShardedSweep(ctx, world, rank, uid[0]); first = False
            sw.set_data(X, y)
            p0, p1, p2, sy = sw.sweep_psi()
            single.set_kernel(1.0, ell); single.set_inducing(Z); single.set_data(X, y)
            f0, f1, f2, fy = single.sweep_psi()
            errs += [float(np.linalg.norm(p2 - f2) / np.linalg.norm(f2)), float(np.linalg.norm(p1 - f1) / max(np.linalg.norm(f1), 1e-300)), abs(p0 - f0), abs(sy - fy) / max(abs(fy), 1e-300)]
    ctx.close(); single.close()
    out[rank] = max(errs)
    dist.destroy_process_group()


def test_two_gpu_ragged_and_empty_shards():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_uneven, args=(2, _free_port(), out), nprocs=2, join=True)
        res = dict(out)
    assert set(res) == {0, 1} and max(res.values()) <= 1e-12, res
