"""N-sharding of the data sweep over ranks.

Psi0 / Psi1 / Psi2 / sum_y2 are sums over independent data points, so the points are cut into contiguous slices, one
per rank (= one process and one ``sgp_ctx`` per GPU); Z and theta are replicated.  Per sweep the ranks exchange ONE
packed buffer ``[Psi2 (M*M) | Psi1 (M*D_out) | Psi0 | sum_y2 | sum_w | n]`` with a sum all-reduce -- inside libsgp that
is ``ncclAllReduce`` on the context's stream (csrc/comm.cu); the helpers here are the host-side part: who owns which
points, and the packed layout (also used by the gloo tests, which run the same exchange on CPU tensors).

The reference has no multi-process code; its only N-scaling device is mini-batching with the
posterior carried as the next prior (helper_functions/gp_helperfunction.jl:137-142), which for fixed theta is the same sum.
"""
import numpy as np


def shard_bounds(N, world, rank):
    """Contiguous slice [lo, hi) of the N points owned by `rank`: sizes differ by at most one, every point has one owner."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank %r / world %r" % (rank, world))
    base, rem = divmod(int(N), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_stats(psi0, psi1, psi2, sum_y2, sum_w, n):
    """The layout of libsgp's resident statistics buffer (csrc/sgp_internal.cuh: stats_dev)."""
    psi2 = np.asarray(psi2, dtype=np.float64)
    psi1 = np.asarray(psi1, dtype=np.float64)
    return np.concatenate([psi2.ravel(order="F"), psi1.ravel(order="F"), np.array([psi0, sum_y2, sum_w, float(n)])])


def unpack_stats(buf, M, D_out=1):
    buf = np.asarray(buf, dtype=np.float64)
    psi2 = buf[:M * M].reshape(M, M, order="F")
    psi1 = buf[M * M:M * M + M * D_out].reshape(M, D_out, order="F")
    psi0, sum_y2, sum_w, n = buf[M * M + M * D_out:M * M + M * D_out + 4]
    return psi0, (psi1[:, 0] if D_out == 1 else psi1), psi2, sum_y2, sum_w, int(round(n))


def tri_col(j, M):
    """Offset of column j in the packed lower triangle (column by column) -- csrc/xchg.cuh: tri_col."""
    return j * M - j * (j - 1) // 2


def pack_stats_lower(psi0, psi1, psi2, sum_y2, sum_w, n):
    """What crosses NVLink (csrc/xchg.cuh): the lower triangle of Psi2 column by column (M (M + 1) / 2 doubles), then Psi1 (M x D_out,
    column-major) and the four scalars; padded with one zero to an even length (the reduce-scatter works on 16-byte pairs)."""
    psi2 = np.asarray(psi2, dtype=np.float64); M = psi2.shape[0]
    cols = [psi2[j:, j] for j in range(M)]
    buf = np.concatenate(cols + [np.asarray(psi1, dtype=np.float64).ravel(order="F"), np.array([psi0, sum_y2, sum_w, float(n)])])
    return np.concatenate([buf, [0.0]]) if buf.size % 2 else buf


def unpack_stats_lower(buf, M, D_out=1):
    """Step 5 of the exchange: packed result -> full symmetric Psi2, Psi1, scalars."""
    buf = np.asarray(buf, dtype=np.float64)
    psi2 = np.empty((M, M))
    for j in range(M):
        c = buf[tri_col(j, M):tri_col(j + 1, M)]
        psi2[j:, j] = c; psi2[j, j:] = c
    tri = M * (M + 1) // 2
    psi1 = buf[tri:tri + M * D_out].reshape(M, D_out, order="F")
    psi0, sum_y2, sum_w, n = buf[tri + M * D_out:tri + M * D_out + 4]
    return psi0, (psi1[:, 0] if D_out == 1 else psi1), psi2, sum_y2, sum_w, int(round(n))


def two_shot_share(n, world, rank):
    """Rank `rank`'s share [lo, hi) of the n packed doubles in the reduce-scatter (pairs of doubles, csrc/xchg.cuh: reduce_scatter)."""
    pairs = (n + 1) // 2
    return 2 * (pairs * rank // world), min(2 * (pairs * (rank + 1) // world), 2 * pairs)


class ShardedSweep:
    """One rank's view of a sharded sweep: owns an SGPContext on `device`, keeps its slice of the data resident and
    attaches the NCCL communicator (the unique id travels over the host's own process group, e.g. torch.distributed)."""

    def __init__(self, ctx, world, rank, uid=None):
        self.ctx, self.world, self.rank = ctx, int(world), int(rank)
        if self.world > 1:
            if uid is None:
                raise ValueError("world > 1 needs the NCCL unique id created on rank 0 (SGPContext.comm_unique_id())")
            ctx.comm_init(self.world, self.rank, uid)

    def set_data(self, X, ybar=None, yvar=None, wts=None):
        """Takes the FULL arrays (every rank holds or can read them) and keeps only this rank's slice on its GPU."""
        lo, hi = shard_bounds(len(X), self.world, self.rank)
        sl = slice(lo, hi)
        self.ctx.set_data(X[sl], None if ybar is None else ybar[sl], None if yvar is None else yvar[sl],
                          None if wts is None else wts[sl])
        return lo, hi

    def sweep_psi(self):
        """Statistics of ALL points on every rank (all-reduced inside sgp_sweep_psi)."""
        return self.ctx.sweep_psi()
