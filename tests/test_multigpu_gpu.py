"""Sharded sweep over two GPUs through the C ABI: N is cut into two slices, one process / sgp_ctx per GPU, the packed
statistics are summed over the ranks inside sgp_sweep_psi -- by the sweep kernel itself through NVLink peer memory (CUDA IPC,
two-shot all-reduce in the kernel's tail), or by one NCCL all-reduce when SGP_COMM_P2P=0; every rank must hold the single-GPU
result (tolerance: the summation order differs, relative Frobenius <= 1e-12) and all ranks the same bits.  Skipped on boxes
with one GPU."""
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
