"""Host-side mirror of the MultiSGP node and of the uncertain-input UniSGP :v rule (enqueue per node, one GPU sweep on the
N-th) against the SUM over nodes of the reference's per-node rules restated in oracle.multisgp / oracle.unisgp.
Pendulum-GPSSM shape (BASELINE.json configs[3]): N = 300 nodes, d_in = 2, D_out = 2, M = 48, srcubature (S = 5)."""
import numpy as np
import pytest

from oracle import cubature as cub, kernels, multisgp, unisgp

pytestmark = pytest.mark.gpu


def fro(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def _setup(N=300, seed=124):
    rng = np.random.default_rng(seed)
    g = np.linspace(-3.0, 3.0, 8); h = np.linspace(-4.0, 4.0, 6)
    Z = np.array([[a, b] for a in g for b in h])                      # 48 inducing points on a grid
    M, D = Z.shape[0], 2
    means = rng.normal(size=(N, 2)) * 1.5
    A = rng.normal(size=(N, 2, 2)) * 0.1
    covs = A @ np.swapaxes(A, 1, 2) + 1e-3 * np.eye(2)
    Y = rng.normal(size=(N, D))
    B = rng.normal(size=(N, D, D)) * 0.05
    Sy = B @ np.swapaxes(B, 1, 2)
    theta = np.array([0.5, 0.8, 1.2])
    kern = lambda t: (kernels.softplus(t[0]), kernels.softplus(t[1:]), 0)
    W = np.array([[90.0, 5.0], [5.0, 110.0]])
    mu_v = rng.normal(size=D * M)
    C = rng.normal(size=(D * M, D * M)) * 0.05
    Sigma_v = C @ C.T + 0.01 * np.eye(D * M)
    var, ell, _ = kern(theta)
    Kinv = np.linalg.inv(kernels.kuu(Z, var, ell, jitter=1e-8))
    return dict(Z=Z, M=M, D=D, N=N, means=means, covs=covs, Y=Y, Sy=Sy, theta=theta, kern=kern, W=W, mu_v=mu_v, Sigma_v=Sigma_v, Kinv=Kinv)


@pytest.mark.parametrize("gaussian_out", [True, False])
def test_multisgp_rules_match_the_sum_of_per_node_rules(gaussian_out):
    from gaussianprocessnode_b200 import nodes as nd
    c = _setup()
    N, M, D = c["N"], c["M"], c["D"]
    ometa = multisgp.MultiSGPMeta(cub.SRCUBATURE, c["Z"], np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), c["Kinv"], c["kern"])
    meta = nd.MultiSGPMeta(cub.SRCUBATURE, c["Z"], np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), None, c["kern"], N=N)
    meta.kuu_jitter = 1e-8
    q_theta = nd.PointMass(c["theta"]); q_w = nd.PointMass(c["W"]); q_v = nd.MvNormalMeanCovariance(c["mu_v"], c["Sigma_v"])
    Sy = c["Sy"] if gaussian_out else [None] * N

    def q_out(n):
        return nd.MvNormalMeanCovariance(c["Y"][n], c["Sy"][n]) if gaussian_out else nd.PointMass(c["Y"][n])

    def q_in(n):
        return nd.MvNormalMeanCovariance(c["means"][n], c["covs"][n])

    # :v -- sum of the per-node (xi_n, Lambda_n)
    o_xi = np.zeros(D * M); o_Lam = np.zeros((D * M, D * M))
    for n in range(N):
        xi, Lam = multisgp.rule_v(c["Y"][n], (c["means"][n], c["covs"][n]), c["W"], c["theta"], ometa)
        o_xi += xi; o_Lam += Lam
    xi = np.zeros(D * M); Lam = np.zeros((D * M, D * M))
    for n in range(N):
        msg = nd.multi_rule_v(q_out(n), q_in(n), q_w, q_theta, meta)
        xi = xi + msg.xi; Lam = Lam + msg.Lam                          # ReactiveMP's prod of weighted-mean/precision Gaussians
        if n < N - 1:
            assert not msg.xi.any() and not msg.Lam.any()
    assert fro(xi, o_xi) < 1e-10 and fro(Lam, o_Lam) < 1e-10

    # :w -- product of N WishartFast(D+2, Psi4_n): nu = N + D + 1, inverse scale = sum_n Psi4_n
    o_P4 = np.zeros((D, D))
    for n in range(N):
        nu, P4 = multisgp.rule_w(c["Y"][n], Sy[n], (c["means"][n], c["covs"][n]), c["mu_v"], c["Sigma_v"], c["theta"], ometa)
        assert nu == D + 2
        o_P4 += P4
    nu_tot = None; P4 = np.zeros((D, D))
    for n in range(N):
        g = nd.multi_rule_w(q_out(n), q_in(n), q_v, q_theta, meta)
        nu_tot = g.nu if nu_tot is None else nu_tot + g.nu - D - 1      # Wishart product rule
        P4 = P4 + g.invS
    assert nu_tot == N * (D + 2) - (N - 1) * (D + 1)
    assert fro(P4, o_P4) < 1e-9, fro(P4, o_P4)                          # includes tr(Kuu^-1 Psi2): conditioning of Kuu + 1e-8 I

    # average energy -- sum_n U_n
    ElogW = float(np.linalg.slogdet(c["W"])[1])
    o_U = sum(multisgp.average_energy(c["Y"][n], Sy[n], (c["means"][n], c["covs"][n]), c["mu_v"], c["Sigma_v"], c["W"], ElogW, c["theta"], ometa)
              for n in range(N))
    U = sum(nd.multi_average_energy(q_out(n), q_in(n), q_v, q_w, q_theta, meta) for n in range(N))
    assert abs(U - o_U) <= 1e-8 * abs(o_U), (U, o_U)

    # :out -- per node mean Psi1_n' mu_v^(d)
    outs = nd.multi_rule_out([q_in(n) for n in range(N)], q_v, q_w, q_theta, meta)
    for n in (0, 7, N - 1):
        om, _ = multisgp.rule_out((c["means"][n], c["covs"][n]), c["mu_v"], c["W"], c["theta"], ometa)
        assert fro(outs[n].m, om) < 1e-10


def test_unisgp_uncertain_v_rule_matches_per_node_fold():
    from gaussianprocessnode_b200 import nodes as nd
    rng = np.random.default_rng(5)
    N, M = 60, 12
    Z = np.linspace(-4, 4, M)
    m = rng.normal(size=N) * 2.0; v = rng.uniform(0.01, 0.3, N); y = rng.normal(size=N)
    theta = np.array([0.4, 0.9]); w = 7.0
    kern = lambda t: (kernels.softplus(t[0]), kernels.softplus(t[1:]), 0)
    method = (cub.GAUSSHERMITE, 21)                                    # GPtest.jl:14 ghcubature(21)
    ometa = unisgp.UniSGPMeta(method, Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), None, kern, np.eye(M), 0, N)
    mu0 = np.zeros(M); S0 = 10.0 * np.eye(M)
    Lam0 = np.linalg.inv(S0); left = (Lam0 @ mu0, Lam0.copy())
    for n in range(N):
        left = unisgp.prod_fold(left, unisgp.rule_v_uncertain(y[n], (m[n], v[n]), w, theta, ometa))
    meta = nd.UniSGPMeta(method, Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), None, kern, None, 0, N)
    marginal = nd.MvNormalMeanCovariance(mu0, S0)
    for n in range(N):
        msg = nd.rule_v_uncertain(nd.PointMass(y[n]), nd.NormalMeanVariance(m[n], v[n]), nd.PointMass(w), nd.PointMass(theta), meta)
        marginal = nd.prod_uncertain(marginal, msg)
    assert fro(marginal.xi, left[0]) < 1e-10 and fro(marginal.Lam, left[1]) < 1e-10
    # the N-th fold is the reference's `prod` (UniSGPnode.jl:62-73): it also delivers mean / covariance and refreshes meta.Uv
    o_mu, o_Sig = unisgp.mean_cov(*left)
    mu, Sig = nd.mean_cov(marginal)
    cond = np.linalg.cond(left[1])
    assert fro(mu, o_mu) < max(1e-9, 50 * cond * 2.2e-16) and fro(Sig, o_Sig) < max(1e-9, 50 * cond * 2.2e-16)
    assert meta.Uv is not None and fro(meta.Uv.T @ meta.Uv, Sig + np.outer(mu, mu)) < 1e-12 and np.allclose(np.tril(meta.Uv, -1), 0.0)
    assert fro(meta.Uv, ometa.Uv) < max(1e-8, 50 * cond * 2.2e-16)


@pytest.mark.parametrize("method", [(cub.GAUSSHERMITE, 21), (cub.SRCUBATURE, 0), (cub.GENUT, 0)])
def test_unisgp_uncertain_w_out_energy_rules_match_per_node_rules(method):
    # @rule UniSGP(:w) / (:out) / @average_energy with q_in Gaussian (UniSGPnode.jl:177-192, 85-93, 290-313): per-NODE results (the reference
    # clamps I1, I2 per node and adds 1e-8 I to every Psi2_n) from sgp_uncertain_node_terms against the oracle's per-node rules
    from gaussianprocessnode_b200 import nodes as nd
    from scipy.linalg import cholesky
    rng = np.random.default_rng(9)
    N, M = 80, 16
    Z = np.linspace(-4, 4, M)
    m = rng.normal(size=N) * 2.0; v = rng.uniform(0.01, 0.3, N); y = rng.normal(size=N); vy = rng.uniform(0.0, 0.2, N)
    theta = np.array([0.4, 0.9])
    kern = lambda t: (kernels.softplus(t[0]), kernels.softplus(t[1:]), 0)
    var, ell, _ = kern(theta)
    jitter = 1e-6
    Kuu = kernels.kuu(Z[:, None], var, ell, jitter=jitter)
    KuuL = np.linalg.cholesky(Kuu)
    mu_v = rng.normal(size=M); C = rng.normal(size=(M, M)) * 0.2
    Sigma_v = C @ C.T + 0.05 * np.eye(M)
    Uv = cholesky(Sigma_v + np.outer(mu_v, mu_v), lower=False)
    ometa = unisgp.UniSGPMeta(method if method[0] == cub.GAUSSHERMITE else method[0], Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), KuuL, kern, Uv, 0, N)
    meta = nd.UniSGPMeta(method if method[0] == cub.GAUSSHERMITE else method[0], Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), None, kern, Uv, 0, N)
    meta.kuu_jitter = jitter
    q_outs = [nd.NormalMeanVariance(y[n], vy[n]) for n in range(N)]
    q_ins = [nd.NormalMeanVariance(m[n], v[n]) for n in range(N)]
    q_v = nd.MvNormalMeanCovariance(mu_v, Sigma_v); q_w = nd.GammaShapeRate(3.0, 0.7); q_theta = nd.PointMass(theta)
    msgs = nd.rule_w_uncertain(q_outs, q_ins, q_v, q_theta, meta)
    outs = nd.rule_out_uncertain(q_ins, q_v, q_w, q_theta, meta)
    U = nd.average_energy_uncertain(q_outs, q_ins, q_v, q_w, q_theta, meta)
    for n in range(N):
        shape, rate = unisgp.rule_w_uncertain(y[n], vy[n], (m[n], v[n]), mu_v, theta, ometa)
        assert msgs[n].a == shape and abs(msgs[n].b - rate) <= 1e-9 * max(abs(rate), 1e-3)
        om, _ = unisgp.rule_out_uncertain((m[n], v[n]), mu_v, 3.0 / 0.7, theta, ometa)
        assert abs(outs[n].m - om) <= 1e-10 * max(abs(om), 1.0)
        oU = unisgp.average_energy_uncertain(y[n], vy[n], (m[n], v[n]), mu_v, (3.0, 0.7), theta, ometa)
        assert abs(U[n] - oU) <= 1e-9 * max(abs(oU), 1.0)
