"""Uncertain-input Psi statistics (sigma-point cloud through the fused weighted sweep; closed-form SE-ARD kernel)
against the oracle.  Tolerance: relative Frobenius <= 1e-10 (north_star)."""
import numpy as np
import pytest

from oracle import batched, cubature as cub

pytestmark = pytest.mark.gpu


def fro(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def ctx():
    from gaussianprocessnode_b200 import SGPContext
    c = SGPContext(0)
    yield c
    c.close()


def _inputs(rng, N, d, scale=0.1):
    mean = rng.normal(size=(N, d)) * 1.5
    A = rng.normal(size=(N, d, d)) * scale
    cov = A @ np.swapaxes(A, 1, 2) + 1e-3 * np.eye(d)
    return mean, cov


def test_pendulum_shape_srcubature_vector_output(ctx):
    # configs[3]: N = 300 MultiSGP nodes, d_in = 2, D_out = 2, M = 48 grid, S = 5 sigma points, W = 100 I
    rng = np.random.default_rng(124)
    g = np.linspace(-3.0, 3.0, 8); h = np.linspace(-4.0, 4.0, 6)
    Z = np.array([[a, b] for a in g for b in h])
    N = 300
    mean, cov = _inputs(rng, N, 2)
    Y = rng.normal(size=(N, 2)); W = 100.0 * np.eye(2)
    ctx.set_kernel(1.3, np.array([1.1, 1.7])); ctx.set_inducing(Z)
    psi0, psi1, psi2, p1n = ctx.sweep_psi_uncertain(cub.SRCUBATURE, mean, cov, R=Y @ W, D_out=2, want_psi1_n=True)
    o0, o1, o2, o1n = batched.psi_stats_uncertain(cub.SRCUBATURE, mean, cov, Z, 1.3, np.array([1.1, 1.7]), YW=Y @ W)
    assert abs(psi0 - o0) < 1e-10 * o0
    assert fro(psi2, o2) < 1e-10 and fro(psi1, o1) < 1e-10 and fro(p1n, o1n) < 1e-10
    # the MultiSGP :v message (MultiSGPnode.jl:290-308 summed): Lambda = kron(W, Psi2), xi = vec(Psi1)
    xi, Lam = batched.multi_v_message(W, psi1, psi2)
    xo, Lo = batched.multi_v_message(W, o1, o2)
    assert fro(xi, xo) < 1e-10 and fro(Lam, Lo) < 1e-10


@pytest.mark.parametrize("method,d,p", [(cub.GENUT, 1, 0), (cub.GENUT, 2, 0), (cub.GAUSSHERMITE, 1, 21), (cub.GAUSSHERMITE, 2, 8),
                                        (cub.SRCUBATURE, 1, 0), (cub.SRCUBATURE, 3, 0)])
def test_methods_scalar_output(ctx, method, d, p):
    rng = np.random.default_rng(10 * method + d)
    N, M = 157, 33
    mean, cov = _inputs(rng, N, d, scale=0.5)
    if method == cub.GENUT:
        cov = np.stack([np.diag(0.3 + rng.random(d)) for _ in range(N)])     # keeps the quirky points at sane distances
    Z = rng.normal(size=(M, d)) * 1.5
    ybar = rng.normal(size=N)
    ell = np.full(d, 1.2)
    ctx.set_kernel(0.9, ell, D=d); ctx.set_inducing(Z)
    psi0, psi1, psi2, p1n = ctx.sweep_psi_uncertain(method, mean, cov, R=ybar, D_out=1, p=max(p, 1), want_psi1_n=True)
    o0, o1, o2, o1n = batched.psi_stats_uncertain(method, mean, cov, Z, 0.9, ell, p=max(p, 1), ybar=ybar)
    tol = 1e-10 if method != cub.GENUT else 1e-9      # GenUT weights are negative / large: cancellation in the sums
    assert abs(psi0 - o0) < tol * max(abs(o0), 1.0)
    assert fro(psi2, o2) < tol and fro(psi1, o1) < tol and fro(p1n, o1n) < tol


@pytest.mark.parametrize("d", [1, 2, 3])
def test_closed_form_se(ctx, d):
    rng = np.random.default_rng(40 + d)
    N, M = 211, 29
    mean, cov = _inputs(rng, N, d, scale=0.4)
    Z = rng.normal(size=(M, d)) * 1.5
    Y = rng.normal(size=(N, 2))
    ell = 0.8 + rng.random(d)
    ctx.set_kernel(1.6, ell, D=d); ctx.set_inducing(Z)
    psi0, psi1, psi2, p1n = ctx.sweep_psi_uncertain(3, mean, cov, R=Y, D_out=2, want_psi1_n=True)
    o0, o1, o2, o1n = batched.psi_stats_closed_form_se(mean, cov, Z, 1.6, ell, YW=Y)
    assert abs(psi0 - o0) < 1e-12 * o0
    assert fro(psi2, o2) < 1e-10 and fro(psi1, o1) < 1e-10 and fro(p1n, o1n) < 1e-10
    if d == 1:    # and the closed form is what high-order Gauss-Hermite converges to
        g0, g1, g2, _ = ctx.sweep_psi_uncertain(cub.GAUSSHERMITE, mean, cov, R=Y, D_out=2, p=48)
        assert fro(g2, psi2) < 1e-9 and fro(g1, psi1) < 1e-9


def test_bad_covariance_is_an_error(ctx):
    from gaussianprocessnode_b200 import SGPError
    ctx.set_kernel(1.0, np.array([1.0, 1.0])); ctx.set_inducing(np.zeros((4, 2)) + np.arange(4)[:, None])
    mean = np.zeros((3, 2)); cov = np.stack([np.eye(2), -np.eye(2), np.eye(2)])
    with pytest.raises(SGPError) as e:
        ctx.sweep_psi_uncertain(cub.SRCUBATURE, mean, cov)
    assert e.value.code == -3
