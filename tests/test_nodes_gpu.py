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


def test_rule_families_may_interleave():
    # ReactiveMP is free to alternate the :v / :w rules and the energy node by node: every interface has its own staging queue
    from gaussianprocessnode_b200 import nodes as nd
    rng = np.random.default_rng(8)
    N, D, M = 40, 2, 10
    X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]); Z = rng.normal(size=(M, D))
    theta = np.array([0.2, 0.4, 0.6]); kern = lambda t: (kernels.softplus(t[0]), kernels.softplus(t[1:]), 0)
    mu0 = np.zeros(M); S0 = 10.0 * np.eye(M); w = 5.0
    q_theta = nd.PointMass(theta); qw = nd.PointMass(w); q_w = nd.GammaShapeRate(2.0, 0.5)
    q_v = nd.MvNormalMeanCovariance(rng.normal(size=M) * 0.1, 0.3 * np.eye(M))
    Uv = np.linalg.cholesky(q_v.S + np.outer(q_v.m, q_v.m)).T

    def run(interleaved):
        meta = nd.UniSGPMeta(None, Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), None, kern, Uv.copy(), 0, N, kuu_jitter=1e-8)
        marg = nd.MvNormalMeanCovariance(mu0, S0); g = None; U = 0.0
        if interleaved:
            for n in range(N):
                marg_n = nd.prod(marg, nd.rule_v(nd.PointMass(y[n]), nd.PointMass(X[n]), qw, q_theta, meta))
                if n < N - 1:
                    marg = marg_n
                else:
                    final = marg_n; meta.Uv = Uv.copy()        # the :w rule / energy of this pass still see the previous sweep's Uv
                g = nd.rule_w(nd.PointMass(y[n]), nd.PointMass(X[n]), q_v, q_theta, meta)
                U += nd.average_energy(nd.PointMass(y[n]), nd.PointMass(X[n]), q_v, q_w, q_theta, meta)
        else:
            for n in range(N):
                marg = nd.prod(marg, nd.rule_v(nd.PointMass(y[n]), nd.PointMass(X[n]), qw, q_theta, meta))
            final = marg; meta.Uv = Uv.copy()
            for n in range(N):
                g = nd.rule_w(nd.PointMass(y[n]), nd.PointMass(X[n]), q_v, q_theta, meta)
            for n in range(N):
                U += nd.average_energy(nd.PointMass(y[n]), nd.PointMass(X[n]), q_v, q_w, q_theta, meta)
        return final, g, U

    fa, ga, Ua = run(False)
    fb, gb, Ub = run(True)
    assert np.array_equal(fa.Lam, fb.Lam) and np.array_equal(fa.xi, fb.xi)
    assert (ga.a, ga.b) == (gb.a, gb.b) and Ua == Ub


@pytest.mark.parametrize("pointmass_in", [True, False])
def test_pointmass_w_energies_with_elementwise_jitter(pointmass_in):
    """@average_energy with q_w::PointMass that adds 1e-8 to EVERY element of an un-jittered K_uu, of Psi1_n and of Psi2_n and calls plain
    `inv` (UniSGPnode.jl:438-458 with q_in PointMass, :390-409 with q_in Gaussian), per node, through the library.  inv(K_uu .+ 1e-8) of the
    un-jittered, near-singular K_uu is defined only up to cond(K_uu) * eps: the test prints the spread between two LAPACK paths (LU `inv`, what
    Julia calls, and a Cholesky solve) and holds the GPU path to the same yardstick."""
    from gaussianprocessnode_b200 import nodes as nd
    from oracle import cubature as cub
    rng = np.random.default_rng(21)
    N, M = 50, 9
    Z = np.linspace(-4, 4, M)                                  # 1-D, spacing 1, lengthscale ~0.7: cond(K_uu) ~ 1e3 -- inv is meaningful
    x = rng.uniform(-4, 4, N); vx = rng.uniform(0.01, 0.2, N); y = rng.normal(size=N); vy = rng.uniform(0.01, 0.2, N)
    theta = np.array([0.5, 0.1]); w = 4.0
    kern = lambda t: (kernels.softplus(t[0]), kernels.softplus(t[1:]), 0)
    var, ell, _ = kern(theta)
    mu_v = rng.normal(size=M) * 0.3; C = rng.normal(size=(M, M)) * 0.1; Sigma_v = C @ C.T + 0.02 * np.eye(M)
    method = (cub.GAUSSHERMITE, 21)
    ometa = unisgp.UniSGPMeta(method, Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), None, kern, np.eye(M), 0, N)
    meta = nd.UniSGPMeta(method, Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), None, kern, None, 0, N)
    q_outs = [nd.NormalMeanVariance(y[n], vy[n]) for n in range(N)]
    q_v = nd.MvNormalMeanCovariance(mu_v, Sigma_v)
    if pointmass_in:
        U = nd.average_energy_gaussout_wpoint(q_outs, [nd.PointMass(x[n]) for n in range(N)], q_v, nd.PointMass(w), nd.PointMass(theta), meta)
        oU = np.array([unisgp.average_energy_gaussout_wpoint(y[n], vy[n], x[n], mu_v, Sigma_v, w, theta, ometa) for n in range(N)])
    else:
        U = nd.average_energy_uncertain_wpoint(q_outs, [nd.NormalMeanVariance(x[n], vx[n]) for n in range(N)], q_v, nd.PointMass(w), nd.PointMass(theta), meta)
        oU = np.array([unisgp.average_energy_uncertain_wpoint(y[n], vy[n], (x[n], vx[n]), mu_v, Sigma_v, w, theta, ometa) for n in range(N)])
    K = kernels.kuu(Z[:, None], var, ell, jitter=0.0) + 1e-8
    lu = np.linalg.inv(K); ch = np.linalg.solve(K, np.eye(M))
    spread = np.linalg.norm(lu - ch) / np.linalg.norm(lu)      # LAPACK vs LAPACK on the same matrix
    condK = np.linalg.cond(K)
    print("cond(K_uu .+ 1e-8) = %.2e, LU-inv vs solve spread = %.2e, GPU vs oracle max rel = %.2e" % (condK, spread, np.max(np.abs(U - oU) / np.abs(oU))))
    tol = max(1e-9, 100 * condK * 2.2e-16)
    assert np.max(np.abs(U - oU) / np.maximum(np.abs(oU), 1.0)) < tol
