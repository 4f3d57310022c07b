"""Driver-side helpers mirrored from helper_functions/gp_helperfunction.jl (CPU only) against the golden chains."""
import numpy as np
import pytest

from gaussianprocessnode_b200 import nodes as nd
from oracle import batched, kernels


def test_split2batch_matches_the_reference_slicing():
    x = list(range(10)); y = list(range(100, 110))
    xb, yb = nd.split2batch((x, y), 4)
    assert xb == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]] and yb == [[100, 101, 102, 103], [104, 105, 106, 107], [108, 109]]
    xb, _ = nd.split2batch((np.arange(10000), np.arange(10000)), 500)      # regression_kin40k.ipynb: 20 batches of 500
    assert len(xb) == 20 and all(len(b) == 500 for b in xb)


def test_smse_reproduces_the_kin40k_notebook_number(kin40k):
    sp = kernels.softplus(kin40k["theta_raw"])
    Z = kin40k["xtrain"][kin40k["xu_ids"]]
    pred = batched.predict_mean(kin40k["xtest"], Z, sp[0], sp[1:], kin40k["mu_v"])
    assert abs(nd.SMSE(kin40k["ytest"], pred) - 0.08343114079545057) < 1e-12       # regression_kin40k.ipynb:315
    assert abs(nd.SMSE(kin40k["ytest"], pred) - batched.smse(kin40k["ytest"], pred)) < 1e-15


def test_error_rate_reproduces_the_banana_notebook_number(banana):
    sp = kernels.softplus(banana["theta_raw"])
    x, lab = banana["x"], banana["label"]
    Z = x[:4000][banana["xu_ids"]]
    xt, yt = x[4000:5300], (lab[4000:5300] > 0).astype(float)                       # -1 -> 0 (classification_banana.ipynb:66-73)
    m = batched.predict_mean(xt, Z, sp[0], sp[1:], banana["mu_v"])
    yhat = (m > 0).astype(float)
    assert nd.num_error(yt, yhat) == 125.0                                           # classification_banana.ipynb:316-317
    assert abs(nd.error_rate(yt, yhat) - 0.09615384615384616) < 1e-15


def test_create_blockmatrix_views():
    A = np.arange(36.0).reshape(6, 6)
    blk = nd.create_blockmatrix(A, 2, 3)
    assert np.array_equal(blk[1][0], A[3:6, 0:3]) and np.shares_memory(blk[0][1], A)


def test_sweep_psi_host_caches_its_marshalling_only_for_the_same_plain_arrays():
    # host logic of the Python mirror (no GPU): the argument tuple of sgp_sweep_psi_host is reused while the caller passes the same
    # C-contiguous float64 arrays, and rebuilt for new arrays, new shapes, converted inputs or a new M / D
    import ctypes
    from gaussianprocessnode_b200 import sgp

    calls = []

    class Lib:
        def sgp_sweep_psi_host(self, *a):
            calls.append(a)
            return 0

    c = sgp.SGPContext.__new__(sgp.SGPContext)
    c.lib = Lib(); c.h = ctypes.c_void_p(1); c.M = 4; c.D = 2; c.N = 0; c._host_call = None
    X = np.zeros((6, 2)); y = np.zeros(6); out = (np.zeros(4), np.zeros((4, 4), order="F"))
    c.sweep_psi_host(X, y, out=out); c.sweep_psi_host(X, y, out=out)
    assert calls[0][2] is calls[1][2] and calls[0][7] is calls[1][7] and calls[0][1] == 6      # the same pointer objects: cached
    X2 = np.ones((6, 2))
    c.sweep_psi_host(X2, y, out=out)
    assert calls[2][2] is not calls[1][2] and c._host_call[0] is X2
    c.sweep_psi_host(X2[:4], y[:4], out=out)                          # views are new objects: rebuilt, N follows
    assert calls[3][1] == 4
    c.sweep_psi_host(np.zeros((6, 2), dtype=np.float32), y, out=out)  # converted temporary: never cached
    assert c._host_call is None
    c.sweep_psi_host(X, y, out=out); c.M = 5
    with __import__("pytest").raises(AssertionError):
        c.sweep_psi_host(X, y, out=out)                               # M changed: the cached outputs no longer fit -> rebuilt and checked


def test_packed_lower_triangle_is_lapack_L_packed_storage():
    # the layout of sgp_sweep_psi_host_packed / sgp_fetch_psi2_packed and of the exchange buffers (csrc/xchg.cuh: tri_col):
    # AP[i + j (2M - j - 1) / 2] = A[i, j] for i >= j (0-based), which is what LAPACK's uplo = 'L' packed routines take
    from gaussianprocessnode_b200.sgp import pack_lower, unpack_lower
    from gaussianprocessnode_b200 import shard
    from scipy.linalg import lapack
    rng = np.random.default_rng(5)
    for M in (1, 2, 7, 33):
        B = rng.normal(size=(M, M + 3)); A = B @ B.T + M * np.eye(M)
        ap = pack_lower(A)
        assert ap.size == M * (M + 1) // 2
        for j in range(M):
            assert shard.tri_col(j, M) == j * (2 * M - j - 1) // 2 + j
            for i in range(j, M):
                assert ap[i + j * (2 * M - j - 1) // 2] == A[i, j]
        assert np.array_equal(unpack_lower(ap, M), A)
        c, info = lapack.dpptrf(M, ap, lower=1)                  # packed Cholesky straight on the vector
        assert info == 0 and np.allclose(np.tril(unpack_lower(c, M)), np.linalg.cholesky(A), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("ncta,nblk,TM", [(148, 4, 128), (148, 2, 64), (148, 8, 128), (148, 16, 128), (148, 20, 128), (148, 32, 128), (37, 4, 128), (3, 2, 64), (1, 1, 128)])
@pytest.mark.parametrize("diag_cost", [None, "0.5", "3.0"])
def test_phase2_plan_covers_every_stripe_once_and_lists_the_partition_slots(ncta, nblk, TM, diag_cost, monkeypatch):
    # host logic of the generate-once sweep's last phase (csrc/sweep.cu: p2_plan_host, through sgp_debug_p2_plan -- no device): every (tile, stripe) of the
    # lower triangle is reduced by exactly one CTA, a CTA never straddles tiles when there are at least as many CTAs as tiles, and the slot list of a tile is
    # the set of CTAs whose share of the cost-weighted (tile, k-step) sequence touches it, in CTA order (slot = cta + tile: a CTA's slots are distinct)
    import ctypes
    from gaussianprocessnode_b200 import _lib, build
    build.build()
    lib = _lib.load()
    if diag_cost is not None:
        monkeypatch.setenv("SGP_SWEEP4_P2_DIAG", diag_cost)
    ntiles = nblk * (nblk + 1) // 2
    stripes = TM // 4
    slab_units = 2504
    w_diag, w_off, w_fixed = 5, 8, 64
    cap = 1 << 20
    out = (ctypes.c_int * cap)(); off = (ctypes.c_int * 4)()
    n = lib.sgp_debug_p2_plan(ncta, ntiles, TM, slab_units, w_diag, w_off, w_fixed, out, cap, off)
    assert n > 0
    a = np.frombuffer(out, dtype=np.int32, count=n)
    cta_off = a[off[0]:off[1]]; items = a[off[1]:off[2]].reshape(-1, 3); tile_off = a[off[2]:off[3]]; slots = a[off[3]:]
    assert len(cta_off) == ncta + 1 and cta_off[0] == 0 and cta_off[-1] == len(items) and np.all(np.diff(cta_off) >= 0)
    cover = np.zeros((ntiles, stripes), dtype=int)
    for c in range(ncta):
        ent = items[cta_off[c]:cta_off[c + 1]]
        if ncta >= ntiles:
            assert len(ent) <= 1
        for t, lo, hi in ent:
            assert 0 <= lo < hi <= stripes and 0 <= t < ntiles
            cover[t, lo:hi] += 1
    assert np.all(cover == 1)
    assert items[0][0] == 0 and items[0][1] == 0 and cta_off[1] >= 1          # CTA 0 starts tile 0, stripe 0: it also finishes the scalars
    # the partition, restated: contiguous shares of the cost sequence [tile 0: w_fixed + w_diag * slab_units | tile 1: ...]
    total = slab_units * (nblk * w_diag + (ntiles - nblk) * w_off) + ntiles * w_fixed
    pos = lambda b: total // ncta * b + total % ncta * b // ncta
    assert len(tile_off) == ntiles + 1 and tile_off[-1] == len(slots)
    pre = 0; t = 0
    for I in range(nblk):
        for J in range(I + 1):
            wt = w_diag if I == J else w_off
            want = []
            for c in range(ncta):
                d0, d1 = pos(c) - pre - w_fixed, pos(c + 1) - pre - w_fixed
                lo = 0 if d0 <= 0 else min(-(-d0 // wt), slab_units)
                hi = 0 if d1 <= 0 else min(-(-d1 // wt), slab_units)
                if lo < hi:
                    want.append(c + t)
            assert list(slots[tile_off[t]:tile_off[t + 1]]) == want and len(want) >= 1
            pre += wt * slab_units + w_fixed; t += 1
    assert len(set(slots.tolist())) == len(slots)
