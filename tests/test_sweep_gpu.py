"""Parity of the CUDA sweep (through the C ABI) against the oracle.  Tolerance from BASELINE.json's north_star:
relative Frobenius error <= 1e-10 on the Psi statistics."""
import numpy as np
import pytest

from oracle import batched, kernels

pytestmark = pytest.mark.gpu
TOL = 1e-10


def fro(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def ctx():
    from gaussianprocessnode_b200 import SGPContext
    c = SGPContext(0)
    yield c
    c.close()


def _check(ctx, X, y, Z, var, ell, yvar=None, w=None, tol=TOL):
    D = Z.shape[1]
    ctx.set_kernel(var, ell, D=D)
    ctx.set_inducing(Z)
    ctx.set_data(X, y, yvar, w)
    psi0, psi1, psi2, sy2 = ctx.sweep_psi()
    o0, o1, o2, oy = batched.psi_stats_point(X, y, Z, var, ell, weights=w, yvar=yvar)
    assert abs(psi0 - o0) <= tol * max(abs(o0), 1.0)
    assert abs(sy2 - oy) <= tol * max(abs(oy), 1.0)
    assert fro(psi1, o1) <= tol, fro(psi1, o1)
    assert fro(psi2, o2) <= tol, fro(psi2, o2)
    assert np.array_equal(psi2, psi2.T)           # exactly symmetric (mirrored, not recomputed)
    return fro(psi2, o2)


def test_toy_regression(ctx, toy):
    # configs[0]: toy 1-D regression, N = 50, M = 20 (GPT_regression.ipynb:765-767), theta = [1, 1]
    X = toy["xtrain_toyregression"][:, None]; y = toy["ytrain_toyregression"]
    Z = np.linspace(-4, 4, 20)[:, None]
    _check(ctx, X, y, Z, 1.0, np.array([1.0]))
    _check(ctx, X, y, Z, 0.0362, np.array([0.5398]))          # the notebook's optimum


def test_kin40k_shape(ctx, kin40k):
    # configs[1]: D = 8, N = 10000, M = 512 (first 512 fixture rows) and the notebook's M = 600
    X, y = kin40k["xtrain"], kin40k["ytrain"]
    sp = kernels.softplus(kin40k["theta_raw"])
    for M in (512, 600):
        Z = X[kin40k["xu_ids"][:M]]
        _check(ctx, X, y, Z, 1.0, np.ones(8))                 # theta_init (regression_kin40k.ipynb:112)
        _check(ctx, X, y, Z, sp[0], sp[1:])                   # fixture optimum
    # the reference's schedule: 20 mini-batches of 500 sum to the full sweep
    Z = X[kin40k["xu_ids"][:512]]
    ctx.set_kernel(sp[0], sp[1:]); ctx.set_inducing(Z)
    acc2 = np.zeros((512, 512)); acc1 = np.zeros(512)
    for b in range(20):
        ctx.set_data(X[500 * b:500 * (b + 1)], y[500 * b:500 * (b + 1)])
        _, p1, p2, _ = ctx.sweep_psi()
        acc1 += p1; acc2 += p2
    ctx.set_data(X, y)
    _, p1, p2, _ = ctx.sweep_psi()
    assert fro(acc2, p2) < 1e-13 and fro(acc1, p1) < 1e-13


def test_banana_shape(ctx, banana):
    # configs[2]: D = 2, N = 4000, M = 64 and the notebook's M = 500; pseudo-targets E[f], Var[f] as in classification
    X = banana["x"][:4000]; lab = (banana["label"][:4000] > 0).astype(float)
    sp = kernels.softplus(banana["theta_raw"])
    rng = np.random.default_rng(1)
    Ef, Vf = batched.probit_moments(rng.normal(size=4000), np.full(4000, 0.5), lab)
    for M in (64, 500):
        Z = X[banana["xu_ids"][:M]]
        _check(ctx, X, Ef, Z, sp[0], sp[1:], yvar=Vf)


@pytest.mark.parametrize("N,D,M", [(1, 1, 1), (31, 3, 7), (33, 2, 65), (1000, 8, 129), (4097, 5, 200), (777, 12, 300), (50, 16, 48)])
def test_ragged_shapes(ctx, N, D, M):
    rng = np.random.default_rng(N + D + M)
    X = rng.normal(size=(N, D)); Z = rng.normal(size=(M, D)); y = rng.normal(size=N)
    ell = 0.7 + rng.random(D) * 2.0
    _check(ctx, X, y, Z, 1.7, ell)


def test_weights_including_negative(ctx):
    rng = np.random.default_rng(5)
    N, D, M = 1500, 2, 48
    X = rng.normal(size=(N, D)); Z = rng.normal(size=(M, D)); y = rng.normal(size=N)
    w = rng.normal(size=N)                                     # sigma-point weights can be negative (GenUT)
    _check(ctx, X, y, Z, 1.0, np.array([1.0, 1.3]), w=w, tol=1e-9)   # cancellation: Psi2 is no longer a sum of positives
    _check(ctx, X, y, Z, 1.0, np.array([1.0, 1.3]), w=np.abs(w))


def test_far_from_origin_inputs(ctx):
    # the expanded-square form is evaluated after centring on the inducing-point mean: offsets must not cost accuracy
    rng = np.random.default_rng(6)
    X = rng.normal(size=(600, 3)) + 1.0e4; Z = rng.normal(size=(40, 3)) + 1.0e4; y = rng.normal(size=600)
    _check(ctx, X, y, Z, 2.0, np.array([1.0, 2.0, 0.5]), tol=1e-9)


@pytest.mark.parametrize("kind", [kernels.MATERN32, kernels.MATERN52])
@pytest.mark.parametrize("N,D,M", [(2000, 8, 300), (333, 2, 64), (1000, 12, 130)])
def test_matern_kernels_in_the_fused_sweep(ctx, kind, N, D, M):
    # KernelFunctions convention r = |(x - z) ./ ell| (the reference only ever imports Matern52Kernel; extension):
    # Matern-3/2 (1 + sqrt3 r) exp(-sqrt3 r), Matern-5/2 (1 + sqrt5 r + 5 r^2 / 3) exp(-sqrt5 r)
    rng = np.random.default_rng(17 + N + kind)
    X = rng.normal(size=(N, D)); Z = np.vstack([X[:M // 2], rng.normal(size=(M - M // 2, D))])   # r = 0 exactly for half of Z
    y = rng.normal(size=N); w = rng.uniform(0.2, 2.0, N)
    ell = 0.8 + rng.random(D) * 2.0
    for wts in (None, w):
        ctx.set_kernel(1.4, ell, D=D, kind=kind); ctx.set_inducing(Z); ctx.set_data(X, y, None, wts)
        psi0, psi1, psi2, sy2 = ctx.sweep_psi()
        o0, o1, o2, oy = batched.psi_stats_point(X, y, Z, 1.4, ell, kind=kind, weights=wts)
        assert abs(psi0 - o0) <= TOL * abs(o0) and fro(psi1, o1) <= TOL and fro(psi2, o2) <= TOL, (fro(psi1, o1), fro(psi2, o2))


@pytest.mark.parametrize("N,D,M", [(3000, 4, 1500), (1500, 8, 2500)])
def test_many_tiles_per_cta(ctx, N, D, M):
    # M = 1500: 78 tiles on 148 CTAs; M = 2500: 210 tiles, i.e. more tiles than CTAs -- a CTA walks through several tiles
    rng = np.random.default_rng(M)
    X = rng.normal(size=(N, D)); Z = rng.normal(size=(M, D)) * 1.5; y = rng.normal(size=N)
    _check(ctx, X, y, Z, 0.9, np.full(D, 1.8))


def test_pinned_buffers_and_segment_clocks(ctx, monkeypatch):
    from gaussianprocessnode_b200 import pinned_empty
    monkeypatch.setenv("SGP_SWEEP_IMPL", "4")          # the record layout checked below is the generate-once kernel's
    rng = np.random.default_rng(21)
    N, D, M = 5000, 8, 256
    X = pinned_empty((N, D)); X[...] = rng.normal(size=(N, D))
    y = pinned_empty((N,)); y[...] = rng.normal(size=N)
    Z = np.array(X[:M])
    psi1 = pinned_empty((M,)); psi2 = pinned_empty((M, M), order="F")
    ctx.set_kernel(1.0, np.full(D, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
    ctx.sweep_debug_clocks()                                   # switches the per-segment clock recording on
    p0, p1, p2, sy = ctx.sweep_psi(out=(psi1, psi2))
    assert p1 is psi1 and p2 is psi2
    o0, o1, o2, oy = batched.psi_stats_point(np.array(X), np.array(y), Z, 1.0, np.full(D, 2.0))
    assert fro(psi2, o2) <= TOL and fro(psi1, o1) <= TOL
    rec = ctx.sweep_debug_clocks()
    rec = rec[rec[:, 0] >= 0]
    seg = rec[rec[:, 2] < 2]                                   # segment records {k-steps, clocks, is_diagonal, cta}; 2 / 3 = generator / waits
    ntiles = 3                                                 # M = 256 -> 2 x 2 blocks of 128 -> 3 lower-triangle tiles
    assert len(seg) >= ntiles and seg[:, 1].min() > 0
    chunks = (N + 31) // 32
    assert seg[:, 0].sum() == 8 * chunks * ntiles              # every (tile, k-step of 4 points) pair is processed exactly once
    gen = rec[rec[:, 2] == 2]
    assert gen[:, 0].sum() == chunks * 2                       # every (row block, chunk) of K_uf is generated exactly once


def test_far_points_underflow_to_zero(ctx):
    X = np.array([[0.0], [1.0e3], [-5.0e4]]); Z = np.array([[0.0], [1.0]]); y = np.ones(3)
    _check(ctx, X, y, Z, 1.0, np.array([1.0]))


def test_determinism(ctx):
    rng = np.random.default_rng(8)
    X = rng.normal(size=(20000, 8)); Z = rng.normal(size=(256, 8)); y = rng.normal(size=20000)
    ctx.set_kernel(1.0, np.full(8, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
    a = ctx.sweep_psi(); b = ctx.sweep_psi()
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1])


def test_synthetic_M1024_exact_on_a_slice_and_properties_at_scale(ctx):
    # configs[4] shape: D = 8, M = 1024, ell = 2, X ~ N(0, I).  Exact check where the oracle finishes in seconds
    # (N = 60k), then size-independent properties at the FULL N = 10M: additivity over a partition and permutation invariance.
    rng = np.random.default_rng(0)
    D, M = 8, 1024
    N = 10_000_000
    X = rng.normal(size=(N, D)); y = np.sin(X @ rng.normal(size=D)) + 0.1 * rng.normal(size=N)
    Z = X[np.random.default_rng(1).choice(N, M, replace=False)]
    ell = np.full(D, 2.0)
    err = _check(ctx, X[:60_000], y[:60_000], Z, 1.0, ell)
    assert err < 1e-12
    ctx.set_kernel(1.0, ell); ctx.set_inducing(Z)
    ctx.set_data(X, y); full = ctx.sweep_psi()
    parts2 = np.zeros((M, M)); parts1 = np.zeros(M); parts0 = 0.0
    for lo, hi in ((0, 3_300_001), (3_300_001, 7_500_000), (7_500_000, N)):
        ctx.set_data(X[lo:hi], y[lo:hi]); p = ctx.sweep_psi()
        parts0 += p[0]; parts1 += p[1]; parts2 += p[2]
    assert fro(parts2, full[2]) < 1e-12 and fro(parts1, full[1]) < 1e-11 and abs(parts0 - full[0]) < 1e-5
    perm = np.random.default_rng(2).permutation(N)
    ctx.set_data(X[perm], y[perm]); q = ctx.sweep_psi()
    assert fro(q[2], full[2]) < 1e-12 and fro(q[1], full[1]) < 1e-11
    # trace identity: tr(Psi2) = sum_n |k_n|^2, checked on a column sample of K
    assert abs(np.trace(full[2]) - np.sum(np.diag(full[2]))) == 0.0
    cols = np.arange(0, M, 128)
    K = kernels.kernel_matrix(X, Z[cols], 1.0, ell)
    np.testing.assert_allclose(np.diag(full[2])[cols], np.sum(K * K, axis=0), rtol=1e-11)
    np.testing.assert_allclose(full[1][cols], K.T @ y, rtol=1e-9, atol=1e-7)


def test_one_call_host_step_equals_set_data_plus_sweep(ctx):
    rng = np.random.default_rng(31)
    N, D, M = 3001, 8, 200
    X = rng.normal(size=(N, D)); Z = rng.normal(size=(M, D)); y = rng.normal(size=N); yv = rng.uniform(0, 0.3, N)
    ctx.set_kernel(1.1, np.full(D, 1.7)); ctx.set_inducing(Z)
    a = ctx.sweep_psi_host(X, y, yv)
    ctx.set_data(X, y, yv); b = ctx.sweep_psi()
    assert a[0] == b[0] and a[3] == b[3] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.array_equal(ctx.sweep_psi()[2], a[2])          # the data stay resident after the one-call step


@pytest.mark.parametrize("N,D,M", [(3001, 8, 200), (10000, 8, 512), (5000, 3, 640), (0, 2, 400)])
def test_host_step_with_packed_psi2_and_upload_beside_the_launch(ctx, N, D, M, monkeypatch):
    # sgp_sweep_psi_host launches the generate-once sweep BESIDE its upload (the kernel waits on the device for the word the copy engine writes
    # after the data) and sgp_sweep_psi_host_packed returns the lower triangle of Psi2 in LAPACK 'L' packed storage: both must give the bits of
    # set_data + sweep_psi, for fresh data on every call (a stale read of the previous batch would show), with pinned and pageable buffers,
    # for the first fused kernel (M <= 384: no device-side wait, packed by a kernel of its own) and with the overlap switched off
    from gaussianprocessnode_b200 import pinned_empty
    from gaussianprocessnode_b200.sgp import pack_lower, unpack_lower
    rng = np.random.default_rng(N + M)
    Z = rng.normal(size=(M, D))
    ctx.set_kernel(1.2, np.full(D, 1.6)); ctx.set_inducing(Z)
    tri = M * (M + 1) // 2
    Xp = pinned_empty((max(N, 1), D))[:N]; yp = pinned_empty((max(N, 1),))[:N]
    buf = pinned_empty((tri + M + 4,))
    for it in range(4):
        X = rng.normal(size=(N, D)); y = rng.normal(size=N); yv = rng.uniform(0, 0.3, N) if it % 2 else None
        Xp[...] = X; yp[...] = y
        if it == 3:
            monkeypatch.setenv("SGP_HOST_OVERLAP", "0")     # upload, then launch (the round-1 order)
        a = ctx.sweep_psi_host(Xp, yp, yv, out=buf, packed=True)
        assert buf[tri + M + 3] == N and np.shares_memory(a[2], buf) and np.shares_memory(a[1], buf)       # [triangle | Psi1 | Psi0, sum_y2, sum_w, n]
        a = (a[0], a[1].copy(), a[2].copy(), a[3])
        f = ctx.sweep_psi_host(X, y, yv)                   # pageable buffers, full square
        ctx.set_data(X, y, yv); b = ctx.sweep_psi()
        assert a[0] == b[0] == f[0] and a[3] == b[3] == f[3]
        assert np.array_equal(a[1], b[1]) and np.array_equal(f[1], b[1]) and np.array_equal(f[2], b[2])
        assert np.array_equal(a[2], pack_lower(b[2])) and np.array_equal(unpack_lower(a[2], M), b[2])
        assert np.array_equal(ctx.fetch_psi2_packed(), a[2])
        if N:
            o0, r1, r2, oy = batched.psi_stats_point(X, y, Z, 1.2, np.full(D, 1.6), yvar=yv)
            assert fro(b[2], r2) <= TOL and fro(b[1], r1) <= TOL and abs(b[3] - oy) <= TOL * abs(oy)


def test_packed_fetch_follows_the_latest_sweep(ctx):
    # sgp_fetch_psi2_packed must return the triangle of the statistics that are resident NOW: a packed copy left by an earlier
    # sgp_sweep_psi_host_packed goes stale with every later sweep (plain, uncertain-input by sigma points, closed form)
    from gaussianprocessnode_b200.sgp import pack_lower
    rng = np.random.default_rng(77)
    N, D, M = 4000, 2, 450
    X = rng.normal(size=(N, D)); y = rng.normal(size=N); Z = rng.normal(size=(M, D))
    ctx.set_kernel(1.1, np.array([1.3, 0.9])); ctx.set_inducing(Z)
    a = ctx.sweep_psi_host(X, y, packed=True)
    assert np.array_equal(ctx.fetch_psi2_packed(), a[2])
    X2 = rng.normal(size=(N // 2, D)); ctx.set_data(X2, y[:N // 2])
    b = ctx.sweep_psi()
    assert np.array_equal(ctx.fetch_psi2_packed(), pack_lower(b[2])) and not np.array_equal(pack_lower(b[2]), a[2])
    ctx.sweep_psi_host(X, y, packed=True)
    mean = rng.normal(size=(300, D)); A_ = rng.normal(size=(300, D, D)) * 0.2; cov = A_ @ np.swapaxes(A_, 1, 2) + 1e-2 * np.eye(D)
    for method in (0, 3):       # srcubature (through the fused sweep), closed-form SE (its own kernels)
        ctx.sweep_psi_host(X, y, packed=True)
        u = ctx.sweep_psi_uncertain(method, mean, cov)
        assert np.array_equal(ctx.fetch_psi2_packed(), pack_lower(u[2]))


@pytest.mark.parametrize("N,D,M,slab_mb,kind", [(20000, 8, 1024, "1", 0), (20000, 3, 300, "0.3", 0), (9000, 2, 100, "0.05", 0), (6000, 8, 300, "0.2", 2)])
def test_many_slabs_ring_wraparound(ctx, N, D, M, slab_mb, kind, monkeypatch):
    # the generate-once kernel cuts N into slabs whose K_uf panel lives in a ring of three L2 panels; tiny panels force dozens of slabs
    # (ring wrap-around, generation / consumption counters, RED accumulation across slabs, clipped last slab) on a small problem
    monkeypatch.setenv("SGP_SWEEP_SLAB_MB", slab_mb)
    monkeypatch.setenv("SGP_SWEEP_IMPL", "4")          # (by default M <= 384 goes to the first fused kernel)
    rng = np.random.default_rng(N + M)
    X = rng.normal(size=(N, D)); Z = rng.normal(size=(M, D)); y = rng.normal(size=N); w = rng.uniform(0.5, 1.5, N)
    ell = 0.9 + rng.random(D) * 1.5
    for wts in (None, w):
        ctx.set_kernel(1.3, ell, D=D, kind=kind); ctx.set_inducing(Z); ctx.set_data(X, y, None, wts)
        psi0, psi1, psi2, sy2 = ctx.sweep_psi()
        o0, o1, o2, oy = batched.psi_stats_point(X, y, Z, 1.3, ell, kind=kind, weights=wts)
        assert abs(psi0 - o0) <= TOL * abs(o0) and fro(psi1, o1) <= TOL and fro(psi2, o2) <= TOL, (fro(psi1, o1), fro(psi2, o2))
    a = ctx.sweep_psi()[2]; b = ctx.sweep_psi()[2]
    assert np.array_equal(a, b)                       # deterministic across launches


def test_first_fused_kernel_still_matches(ctx, monkeypatch):
    # SGP_SWEEP_IMPL=3: the first fused kernel (K_uf regenerated in shared memory per tile), kept as the fallback path
    monkeypatch.setenv("SGP_SWEEP_IMPL", "3")
    rng = np.random.default_rng(77)
    X = rng.normal(size=(5000, 8)); Z = rng.normal(size=(300, 8)); y = rng.normal(size=5000)
    _check(ctx, X, y, Z, 1.1, np.full(8, 1.7))


@pytest.mark.parametrize("M", [48, 512])
def test_empty_data_set_gives_zero_statistics(ctx, M):
    # a rank whose shard is empty still calls the sweep: zeros, not an error (and the following M x M calls see zero statistics)
    rng = np.random.default_rng(0)
    Z = rng.normal(size=(M, 3))
    ctx.set_kernel(1.0, np.full(3, 1.0)); ctx.set_inducing(Z); ctx.set_data(np.zeros((0, 3)), np.zeros(0))
    p0, p1, p2, sy = ctx.sweep_psi()
    assert p0 == 0.0 and sy == 0.0 and not p1.any() and not p2.any()
    mu, Sig, Uv = ctx.posterior_v(np.zeros(M), np.eye(M) / 4.0, 10.0)
    assert np.allclose(mu, 0.0) and np.allclose(Sig, 4.0 * np.eye(M), rtol=1e-13, atol=1e-13)


def test_single_and_double_slab_policy_matches_forced_ring(ctx, monkeypatch):
    # small sweeps run as one (<= 64 MB of K_uf) or two (<= 96 MB) slabs; the same numbers as the forced 3-panel ring of small panels
    rng = np.random.default_rng(4)
    D, M = 8, 512
    Z = rng.normal(size=(M, D))
    for N in (9000, 19000):                        # 38 MB -> one slab, 80 MB -> two slabs
        X = rng.normal(size=(N, D)); y = np.sin(X[:, 1])
        ctx.set_kernel(1.0, np.full(D, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
        a = ctx.sweep_psi()
        monkeypatch.setenv("SGP_SWEEP_SLAB_MB", "7")
        b = ctx.sweep_psi()
        monkeypatch.delenv("SGP_SWEEP_SLAB_MB")
        assert fro(a[2], b[2]) < 1e-14 and fro(a[1], b[1]) < 1e-14 and a[0] == b[0]
        o = batched.psi_stats_point(X, y, Z, 1.0, np.full(D, 2.0))
        assert fro(a[2], o[2]) < 1e-13 and fro(a[1], o[1]) < 1e-13
