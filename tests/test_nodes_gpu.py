"""The host-side mirror of the node interface (enqueue in the per-point rules, flush in the N-th prod) produces the
same marginals / messages as the reference's per-point schedule restated in oracle.unisgp."""
import numpy as np
import pytest

from oracle import batched, kernels, unisgp

pytestmark = pytest.mark.gpu


def fro(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("classification", [False, True])
def test_unisgp_pass_matches_per_point_reference_schedule(classification):
    from gaussianprocessnode_b200 import nodes as nd
    rng = np.random.default_rng(3)
    N, D, M = 90, 3, 14
    X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]) + 0.05 * rng.normal(size=N); yv = rng.random(N) * 0.2 if classification else np.zeros(N)
    Z = rng.normal(size=(M, D)) * 1.3
    theta = np.array([0.3, 0.1, 0.5, 0.9])
    kern = lambda t: (kernels.softplus(t[0]), kernels.softplus(t[1:]), 0)
    w = 20.0
    mu0 = np.zeros(M); S0 = 50.0 * np.eye(M)
    # ---- reference schedule (oracle) ----
    var, ell, _ = kern(theta)
    Lo = np.linalg.cholesky(kernels.kuu(Z, var, ell, jitter=1e-8))
    ometa = unisgp.UniSGPMeta(None, Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), Lo, kern, np.eye(M), 0, N)
    o_mu, o_Sig, o_xi, o_Lam = unisgp.sweep_v_pointmass(X, y, w, theta, ometa, mu0, S0)
    o_rate = sum(unisgp.rule_w_pointmass(y[n], yv[n], X[n], o_mu, theta, ometa)[1] for n in range(N))
    q_w = (3.0, 0.2)
    o_U = sum(unisgp.average_energy_pointmass(y[n], yv[n], X[n], o_mu, q_w, theta, ometa) for n in range(N))
    # ---- mirror ----
    meta = nd.UniSGPMeta(None, Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), None, kern, None, 0, N, kuu_jitter=1e-8)
    q_theta = nd.PointMass(theta); qw = nd.PointMass(w)
    marginal = nd.MvNormalMeanCovariance(mu0, S0)
    for n in range(N):
        q_out = nd.NormalMeanVariance(y[n], yv[n]) if classification else nd.PointMass(y[n])
        msg = nd.rule_v(q_out, nd.PointMass(X[n]), qw, q_theta, meta)
        assert isinstance(msg, nd.BufferUniSGP)
        marginal = nd.prod(marginal, msg)
    assert meta.counter == 0
    mu, Sig = nd.mean_cov(marginal)
    assert fro(marginal.Lam, o_Lam) < 1e-10 and fro(marginal.xi, o_xi) < 1e-10
    assert fro(mu, o_mu) < 1e-9 and fro(Sig, o_Sig) < 1e-9 and fro(meta.Uv, ometa.Uv) < 1e-9
    q_v = nd.MvNormalMeanCovariance(o_mu, o_Sig)
    meta.Uv = ometa.Uv                      # same lagged Uv on both sides
    g = None
    for n in range(N):
        q_out = nd.NormalMeanVariance(y[n], yv[n]) if classification else nd.PointMass(y[n])
        g = nd.rule_w(q_out, nd.PointMass(X[n]), q_v, q_theta, meta)
        if n < N - 1:
            assert (g.a, g.b) == (1.0, 0.0)
    assert g.a == 1.0 + N / 2 and abs(g.b - o_rate) < 1e-9 * abs(o_rate)
    U = 0.0
    for n in range(N):
        q_out = nd.NormalMeanVariance(y[n], yv[n]) if classification else nd.PointMass(y[n])
        U += nd.average_energy(q_out, nd.PointMass(X[n]), q_v, nd.GammaShapeRate(*q_w), q_theta, meta)
    assert abs(U - o_U) < 1e-8 * abs(o_U)
    if not classification:
        # q_out, q_in, q_w all PointMass (UniSGPnode.jl:411-436): E[ln w] = ln w, chol(Sigma_v + mu mu').U recomputed by the rule
        o_Up = sum(unisgp.average_energy_pointmass_wpoint(y[n], X[n], o_mu, o_Sig, w, theta, ometa) for n in range(N))
        Up = 0.0
        for n in range(N):
            Up += nd.average_energy(nd.PointMass(y[n]), nd.PointMass(X[n]), q_v, qw, q_theta, meta)
        assert abs(Up - o_Up) < 1e-8 * abs(o_Up)
    out = nd.rule_out(nd.PointMass(X[0]), q_v, qw, q_theta, meta)
    ref_m, _ = unisgp.rule_out_pointmass(X[0], o_mu, w, theta, ometa)
    assert abs(out.m - ref_m) < 1e-10 * max(abs(ref_m), 1.0) and out.w == w
