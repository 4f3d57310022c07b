"""Driver-side helpers mirrored from helper_functions/gp_helperfunction.jl (CPU only) against the golden chains."""
import numpy as np

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
