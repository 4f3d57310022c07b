"""Oracle of the theta objective: the batched identity equals the reference's per-point loop, and the analytic gradient
equals central differences of it (what ForwardDiff returns up to rounding)."""
import numpy as np
import pytest

from oracle import kernels, theta


def _case(seed, N, D, M):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, D)); Z = rng.normal(size=(M, D)) * 1.2; y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=N)
    v = rng.normal(size=M)
    C = rng.normal(size=(M, M)) * 0.3
    Uv = np.linalg.cholesky(C @ C.T + 0.5 * np.eye(M)).T
    return X, y, Z, v, Uv


@pytest.mark.parametrize("kind", [kernels.SE, kernels.MATERN32, kernels.MATERN52])
def test_batched_objective_equals_per_point_loop_and_gradient_equals_finite_differences(kind):
    X, y, Z, v, Uv = _case(1, 120, 3, 15)
    var, ell, w, jit = 1.3, np.array([0.9, 1.4, 2.0]), 7.0, 1e-6
    F, dvar, dell = theta.objective_and_gradient(var, ell, y, X, v, Uv, w, Z, kind, jit)
    F_loop = theta.neg_log_backwardmess_fast(var, ell, y, X, v, Uv, w, Z, kind, jit)
    assert abs(F - F_loop) <= 1e-10 * max(abs(F_loop), 1.0)
    f = lambda va, el: theta.neg_log_backwardmess_fast(va, el, y, X, v, Uv, w, Z, kind, jit)
    eps = 1e-5
    fd = (f(var + eps, ell) - f(var - eps, ell)) / (2 * eps)
    assert abs(dvar - fd) <= 1e-6 * max(abs(fd), 1.0), (dvar, fd)
    for d in range(3):
        e = np.zeros(3); e[d] = eps
        fd = (f(var, ell + e) - f(var, ell - e)) / (2 * eps)
        assert abs(dell[d] - fd) <= 1e-6 * max(abs(fd), 1.0), (d, dell[d], fd)
