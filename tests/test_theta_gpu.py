"""The theta step: collapsed objective and its exact gradient on the GPU (sgp_theta_objective) against the
oracle's per-point restatement of derivative_helper.jl:23-39 and its analytic gradient.  Tolerances: value 1e-9 relative
(it contains tr(K_uu^-1 Psi2): conditioning of K_uu), gradient 1e-8 relative to the gradient's norm."""
import numpy as np
import pytest

from oracle import kernels, theta as otheta

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from gaussianprocessnode_b200 import SGPContext
    c = SGPContext(0)
    yield c
    c.close()


def _case(seed, N, D, M):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, D)); Z = rng.normal(size=(M, D)) * 1.2; y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=N)
    v = rng.normal(size=M)
    C = rng.normal(size=(M, M)) * 0.3
    Uv = np.linalg.cholesky(C @ C.T + 0.5 * np.eye(M)).T
    return X, y, Z, v, Uv


@pytest.mark.parametrize("kind", [kernels.SE, kernels.MATERN32, kernels.MATERN52])
@pytest.mark.parametrize("N,D,M", [(500, 8, 200), (333, 2, 48), (40000, 3, 100)])
def test_objective_and_gradient(ctx, kind, N, D, M):
    X, y, Z, v, Uv = _case(N + kind, N, D, M)
    var, ell, w, jit = 1.3, 0.8 + np.arange(D) * 0.15, 11.0, 1e-6
    ctx.set_kernel(var, ell, D=D, kind=kind); ctx.set_inducing(Z); ctx.set_data(X, y)
    F, dvar, dell = ctx.theta_objective(v, Uv, w, jit)
    oF, odvar, odell = otheta.objective_and_gradient(var, ell, y, X, v, Uv, w, Z, kind, jit)
    assert abs(F - oF) <= 1e-9 * abs(oF), (F, oF)
    g = np.concatenate([[dvar], dell]); og = np.concatenate([[odvar], odell])
    assert np.linalg.norm(g - og) <= 1e-8 * np.linalg.norm(og), (g, og)
    assert ctx.theta_objective(v, Uv, w, jit, grad=False) == F          # value-only path, deterministic


def test_reference_signature_and_softplus_chain_rule(ctx):
    # the kin40k driver's call (regression_kin40k.ipynb:214-222): mini-batch of 500, theta raw with softplus
    from gaussianprocessnode_b200 import theta as th
    X, y, Z, v, Uv = _case(7, 500, 8, 120)
    raw = np.array([0.2, 0.5, 0.1, 0.9, 1.3, 0.4, 0.7, 1.1, 0.3])
    kern = lambda t: (kernels.softplus(t[0]), kernels.softplus(t[1:]), 0)
    F = th.neg_log_backwardmess_fast(raw, y_data=y, x_data=X, v=v, Uv=Uv, w=1.0e4, kernel=kern, Xu=Z, ctx=ctx)
    oF = otheta.neg_log_backwardmess_fast(kernels.softplus(raw[0]), kernels.softplus(raw[1:]), y, X, v, Uv, 1.0e4, Z)
    assert abs(F - oF) <= 1e-8 * abs(oF)
    grad = np.zeros(9)
    th.grad_llh_new(grad, raw, y_data=y, x_data=X, v=v, Uv=Uv, w=1.0e4, kernel=kern, Xu=Z, chunk_size=4, ctx=ctx)
    _, odv, odl = otheta.objective_and_gradient(kernels.softplus(raw[0]), kernels.softplus(raw[1:]), y, X, v, Uv, 1.0e4, Z)
    sig = 1.0 / (1.0 + np.exp(-raw))
    og = np.concatenate([[odv], odl]) * sig
    assert np.linalg.norm(grad - og) <= 1e-7 * np.linalg.norm(og)
