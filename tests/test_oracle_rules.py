"""GPtest.jl's closed-form ground truths, re-expressed with seeded inputs, plus the per-point == batched proofs.

Reference test file: /root/reference/GPtest.jl (unseeded; it holds formulas, not numbers).  Inputs mirror it:
Xu = 1:10 (univariate), 5x5 grid (multivariate), kernel = theta1 * SE(lengthscale theta2), theta = [1, 1]."""
import numpy as np
import pytest
from scipy.special import digamma

from oracle import batched, cubature as cub, kernels, multisgp, unisgp

THETA = np.array([1.0, 1.0])
KERN = lambda t: (t[0], t[1], kernels.SE)  # GPtest.jl:21 (raw theta, no softplus)


def _uni_meta(rng, method=(cub.GAUSSHERMITE, 21)):
    Xu = np.arange(1.0, 11.0)[:, None]
    Kuu = kernels.kuu(Xu, 1.0, 1.0)
    mu_v = np.sin(rng.random(10)); Sig_v = np.eye(10)
    Rv = Sig_v + np.outer(mu_v, mu_v)
    meta = unisgp.UniSGPMeta(method, Xu, np.ones((1, 1)), np.zeros((10, 1)), np.zeros((10, 10)),
                             np.linalg.cholesky(Kuu), KERN, np.linalg.cholesky(Rv).T, 0, 10)
    return meta, Xu, Kuu, mu_v, Sig_v, Rv


def test_uni_cubature_vs_monte_carlo_and_high_order():
    # GPtest.jl:127-143 (5000 MC samples; atol 1e-4 / 0.05 / 0.05)
    rng = np.random.default_rng(0)
    meta, Xu, *_ = _uni_meta(rng)
    psi0, psi1, psi2 = unisgp.kernel_expectations(meta, THETA, 0.0, 1.0)
    xs = rng.normal(size=5000)
    K = kernels.kernel_matrix(xs[:, None], Xu, 1.0, 1.0)
    assert abs(psi0 - 1.0) < 1e-4
    assert np.abs(K.mean(0) - psi1).max() < 0.05
    assert np.abs(K.T @ K / 5000 - psi2).max() < 0.05
    meta.method = (cub.GAUSSHERMITE, 60)
    _, p1, p2 = unisgp.kernel_expectations(meta, THETA, 0.0, 1.0)
    assert np.abs(p1 - psi1).max() < 1e-6 and np.abs(p2 - psi2).max() < 1e-5   # 21-point rule vs 60-point rule
    # closed-form SE (extension) agrees with high-order GH
    _, c1, c2, _ = batched.psi_stats_closed_form_se(np.zeros((1, 1)), np.ones((1, 1, 1)), Xu, 1.0, 1.0)
    assert np.abs(c1 - p1).max() < 1e-13 and np.abs(c2 - p2).max() < 1e-13


def test_uni_rule_v():
    # GPtest.jl:183-217
    rng = np.random.default_rng(1)
    meta, Xu, *_ = _uni_meta(rng)
    w = 1.0  # mean(GammaShapeRate(1,1))
    _, psi1, psi2 = unisgp.kernel_expectations(meta, THETA, 0.0, 1.0)
    msg = unisgp.rule_v_uncertain(1.0, (0.0, 1.0), w, THETA, meta)
    mean, cov = unisgp.mean_cov(msg.xi, msg.Lam)
    np.testing.assert_allclose(mean, np.linalg.inv(psi2 + 1e-8 * np.eye(10)) @ psi1 * 1.0, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(cov, np.linalg.inv(w * (psi2 + 1e-8 * np.eye(10))), rtol=1e-6)
    k = kernels.kernel_matrix(np.array([[1.0]]), Xu, 1.0, 1.0)[0]
    msg = unisgp.rule_v_pointmass(2.0, 1.0, w, THETA, meta)
    np.testing.assert_allclose(msg.xi, 2.0 * w * k, rtol=1e-15)
    np.testing.assert_allclose(msg.Lam, w * np.outer(k, k), rtol=1e-15)
    assert msg.Lam is meta.Psi2          # aliasing quirk (UniSGPnode.jl:155-157)


def test_uni_rule_w_and_energy():
    # GPtest.jl:220-254, 294-349
    rng = np.random.default_rng(2)
    meta, Xu, Kuu, mu_v, Sig_v, Rv = _uni_meta(rng)
    Kinv = np.linalg.inv(Kuu)
    psi0, psi1, psi2 = unisgp.kernel_expectations(meta, THETA, 0.0, 1.0)
    I1 = psi0 - np.trace(Kinv @ psi2); I2 = 1.0 + 4.0 - 2 * 1.0 * (psi1 @ mu_v) + np.trace(Rv @ psi2)
    shape, rate = unisgp.rule_w_uncertain(1.0, 4.0, (0.0, 1.0), mu_v, THETA, meta)   # q_out = Normal(1, 2): var 4
    assert shape == 1.5 and abs(rate - 0.5 * (I1 + I2)) < 1e-5
    k = kernels.kernel_matrix(np.array([[1.0]]), Xu, 1.0, 1.0)[0]
    I1 = 1.0 - k @ Kinv @ k; I2 = 4.0 - 4.0 * (k @ mu_v) + k @ Rv @ k
    shape, rate = unisgp.rule_w_pointmass(2.0, 0.0, 1.0, mu_v, THETA, meta)
    assert shape == 1.5 and abs(rate - 0.5 * (I1 + I2)) < 1e-5
    I2c = 1.0 + 4.0 - 2.0 * (k @ mu_v) + k @ Rv @ k
    assert abs(unisgp.rule_w_pointmass(1.0, 4.0, 1.0, mu_v, THETA, meta)[1] - 0.5 * (I1 + I2c)) < 1e-5
    # energies (q_w = GammaShapeRate(1,1))
    E_logw = digamma(1.0)
    U = unisgp.average_energy_pointmass(2.0, 0.0, 1.0, mu_v, (1.0, 1.0), THETA, meta)
    assert abs(U - 0.5 * (I1 - E_logw + np.log(2 * np.pi) + I2)) < 1e-6
    U = unisgp.average_energy_pointmass(1.0, 4.0, 1.0, mu_v, (1.0, 1.0), THETA, meta)
    assert abs(U - 0.5 * (I1 - E_logw + np.log(2 * np.pi) + I2c)) < 1e-6
    U = unisgp.average_energy_pointmass_wpoint(2.0, 1.0, mu_v, Sig_v, 2.5, THETA, meta)
    assert abs(U - 0.5 * (I1 * 2.5 - np.log(2.5) + np.log(2 * np.pi) + I2 * 2.5)) < 1e-6
    U = unisgp.average_energy_gaussout_wpoint(1.0, 4.0, 1.0, mu_v, Sig_v, 2.5, THETA, meta)
    assert abs(U - 0.5 * (I1 * 2.5 - np.log(2.5) + np.log(2 * np.pi) + I2c * 2.5)) < 1e-5


@pytest.mark.parametrize("D,M,N", [(1, 6, 50), (3, 17, 64), (8, 40, 100)])
def test_uni_per_point_schedule_equals_batched(D, M, N):
    """N rule invocations + prod folds (UniSGPnode.jl:144-158, 62-73, 196-238, 337-387) == one batched sweep."""
    rng = np.random.default_rng(10 + D)
    Z = rng.normal(size=(M, D)); X = rng.normal(size=(N, D)); y = rng.normal(size=N); yv = rng.random(N)
    if D == 1:
        Z = np.linspace(-2.5, 2.5, M)[:, None]      # random 1-D inducing points make K_uu singular to working precision
    th = np.concatenate([[1.3], 0.8 + rng.random(D)])
    kern = lambda t: (t[0], t[1:], kernels.SE)
    L = np.linalg.cholesky(kernels.kuu(Z, th[0], th[1:], jitter=1e-8))
    meta = unisgp.UniSGPMeta(None, Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), L, kern, np.eye(M), 0, N)
    w = 7.0; mu0 = rng.normal(size=M) * 0.1; S0 = 50.0 * np.eye(M)
    mu, Sig, xi, Lam = unisgp.sweep_v_pointmass(X, y, w, th, meta, mu0, S0)
    p0, p1, p2, sy2 = batched.psi_stats_point(X, y, Z, th[0], th[1:], yvar=yv)
    mu_b, Sig_b, Uv_b, Lam_b, xi_b = batched.posterior_v(mu0 / 50.0, np.eye(M) / 50.0, w, p1, p2)
    fro = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert fro(Lam, Lam_b) < 1e-13 and fro(xi, xi_b) < 1e-13
    assert fro(mu, mu_b) < 1e-9 and fro(Sig, Sig_b) < 1e-9 and fro(meta.Uv, Uv_b) < 1e-9
    rates = sum(unisgp.rule_w_pointmass(y[n], yv[n], X[n], mu, th, meta)[1] for n in range(N))
    s1, s2 = batched.w_terms(p0, p1, p2, sy2, L, mu, meta.Uv)
    assert abs(rates - 0.5 * (s1 + s2)) < 1e-10 * abs(rates)
    q_w = (3.0, 0.4)
    U = sum(unisgp.average_energy_pointmass(y[n], yv[n], X[n], mu, q_w, th, meta) for n in range(N))
    w_bar, E_logw = unisgp.gamma_stats(*q_w)
    assert abs(U - batched.energy_sum(N, s1, s2, w_bar, E_logw)) < 1e-10 * abs(U)


def _multi_setup(rng):
    Xu = np.array([[i, j] for j in range(1, 6) for i in range(1, 6)], dtype=float)   # GPtest.jl:20
    Kinv = batched.cholinv(kernels.kuu(Xu, 1.0, 1.0, jitter=1e-12))
    meta = multisgp.MultiSGPMeta((cub.SRCUBATURE, 0), Xu, np.ones((1, 1)), np.zeros((25, 1)), np.zeros((25, 25)), Kinv, KERN)
    mu_v = np.sin(rng.random(50)); Sig_v = np.eye(50)
    return meta, Xu, Kinv, mu_v, Sig_v


def test_multi_rules_against_kron_ground_truths():
    # GPtest.jl:352-539
    rng = np.random.default_rng(3)
    meta, Xu, Kinv, mu_v, Sig_v = _multi_setup(rng)
    C = np.eye(2)
    mu_y = np.array([0.5, 1.4]); Sig_y = np.eye(2)
    q_in = (np.array([1.0, 2.7]), np.eye(2))
    W, E_logW = batched.wishart_stats(10.0, 50.0 * np.eye(2))
    Rv = Sig_v + np.outer(mu_v, mu_v)
    psi0, psi1, psi2 = multisgp.kernel_expectations(meta, THETA, *q_in)
    assert psi0 == 1.0                                    # GPtest.jl:380: weights sum to one exactly
    xs = rng.multivariate_normal(q_in[0], q_in[1], size=10000)
    K = kernels.kernel_matrix(xs, Xu, 1.0, 1.0)
    assert np.abs(K.mean(0) - psi1).max() < 0.08 and np.abs(K.T @ K / 10000 - psi2).max() < 0.3
    # :out
    mean, prec = multisgp.rule_out(q_in, mu_v, W, THETA, meta)
    np.testing.assert_allclose(mean, np.kron(C, psi1[None, :]) @ mu_v, rtol=1e-13)
    # :v
    xi, Lam = multisgp.rule_v(mu_y, q_in, W, THETA, meta)
    np.testing.assert_allclose(Lam, np.kron(W, psi2), rtol=1e-15)
    np.testing.assert_allclose(xi, np.kron(C, psi1[None, :]).T @ W @ mu_y, rtol=1e-13)
    # :w  (Psi_4 = E[kron(C,k') R_v kron(C,k)] by the same cubature)
    pts, wts = cub.srcubature(*q_in)
    Psi4 = sum(wt * np.kron(C, kernels.kernel_matrix(pt[None], Xu, 1.0, 1.0)) @ Rv @ np.kron(C, kernels.kernel_matrix(pt[None], Xu, 1.0, 1.0)).T
               for pt, wt in zip(pts, wts))
    I1 = np.kron(C, psi0 - np.trace(Kinv @ psi2))
    T1 = np.kron(C, psi1[None, :])
    I2 = np.outer(mu_y, mu_y) + Sig_y - np.outer(mu_y, mu_v) @ T1.T - T1 @ np.outer(mu_v, mu_y) + Psi4
    nu, invS = multisgp.rule_w(mu_y, Sig_y, q_in, mu_v, Sig_v, THETA, meta)
    assert nu == 4
    np.testing.assert_allclose(invS, I1 + I2, rtol=1e-10)
    # energy (GPtest.jl:509-538, atol 1e-2 there)
    U = multisgp.average_energy(mu_y, Sig_y, q_in, mu_v, Sig_v, W, E_logW, THETA, meta)
    U_gt = 0.5 * np.trace(W @ I1) + 0.5 * 2 * np.log(2 * np.pi) - 0.5 * E_logW + 0.5 * np.trace(W @ I2)
    assert abs(U - U_gt) < 1e-8 * abs(U_gt)


def test_multi_per_point_equals_batched():
    rng = np.random.default_rng(4)
    meta, Xu, Kinv, mu_v, Sig_v = _multi_setup(rng)
    N, D, M = 30, 2, 25
    means = rng.normal(size=(N, 2)) * 1.5 + 3.0
    A = rng.normal(size=(N, 2, 2)) * 0.2
    covs = A @ np.swapaxes(A, 1, 2) + 0.05 * np.eye(2)
    Y = rng.normal(size=(N, D))
    W = np.array([[90.0, 5.0], [5.0, 110.0]])
    xi = np.zeros(D * M); Lam = np.zeros((D * M, D * M))
    for n in range(N):
        a, b = multisgp.rule_v(Y[n], (means[n], covs[n]), W, THETA, meta)
        xi += a; Lam += b
    p0, p1, p2, p1n = batched.psi_stats_uncertain(cub.SRCUBATURE, means, covs, Xu, 1.0, 1.0, YW=Y @ W)
    xi_b, Lam_b = batched.multi_v_message(W, p1, p2)
    assert np.linalg.norm(xi - xi_b) / np.linalg.norm(xi) < 1e-13
    assert np.linalg.norm(Lam - Lam_b) / np.linalg.norm(Lam) < 1e-13
    assert abs(p0 - N) < 1e-12


def test_sigma_point_rules_moments():
    rng = np.random.default_rng(5)
    m = rng.normal(size=3); A = rng.normal(size=(3, 3)); P = A @ A.T + np.eye(3)
    for pts, wts in (cub.srcubature(m, P), cub.ghcubature(4, m, P)):
        assert abs(wts.sum() - 1) < 1e-14
        np.testing.assert_allclose(wts @ pts, m, atol=1e-13)
        d = pts - m
        np.testing.assert_allclose((d * wts[:, None]).T @ d, P, atol=1e-12)
    pts, wts = cub.gen_unscented_uni(0.3, 1.0)          # V = 1: the classic (2/3, 1/6, 1/6) at m, m -/+ sqrt(3)
    np.testing.assert_allclose(pts[:, 0], [0.3, 0.3 - np.sqrt(3), 0.3 + np.sqrt(3)], rtol=1e-15)
    np.testing.assert_allclose(wts, [2 / 3, 1 / 6, 1 / 6], rtol=1e-14)
    pts, wts = cub.gen_unscented_uni(0.0, 4.0)          # quirk: m -/+ sqrt(3)/sqrt(V)  (ut_approx.jl:116-126)
    np.testing.assert_allclose(pts[1:, 0], [-np.sqrt(3) / 2, np.sqrt(3) / 2], rtol=1e-15)
    pts, wts = cub.gen_unscented_multi(np.zeros(2), np.eye(2))
    assert abs(wts.sum() - 1) < 1e-14 and pts.shape == (5, 2)


def test_closed_form_se_vs_tensor_gauss_hermite_2d():
    rng = np.random.default_rng(6)
    Z = rng.normal(size=(7, 2)); m = rng.normal(size=(3, 2)); A = rng.normal(size=(3, 2, 2)) * 0.4
    S = A @ np.swapaxes(A, 1, 2) + 0.02 * np.eye(2)
    ell = np.array([0.9, 1.4])
    _, c1, c2, _ = batched.psi_stats_closed_form_se(m, S, Z, 1.7, ell)
    _, g1, g2, _ = batched.psi_stats_uncertain(cub.GAUSSHERMITE, m, S, Z, 1.7, ell, p=40)
    assert np.abs(c1 - g1).max() < 1e-12 and np.abs(c2 - g2).max() < 1e-9    # limited by the 40x40 rule


def test_free_energy_assembly_matches_per_node_sum():
    rng = np.random.default_rng(7)
    N, M = 25, 8
    Z = np.linspace(-3, 3, M)[:, None]; X = rng.normal(size=(N, 1)) * 2; y = np.sin(X[:, 0])
    th = np.array([1.1, 0.7]); kern = lambda t: (t[0], t[1], kernels.SE)
    L = np.linalg.cholesky(kernels.kuu(Z, 1.1, 0.7, jitter=1e-8))
    p0, p1, p2, sy2 = batched.psi_stats_point(X, y, Z, 1.1, 0.7)
    mu, Sig, Uv, _, _ = batched.posterior_v(np.zeros(M), np.eye(M), 2.0, p1, p2)
    s1, s2 = batched.w_terms(p0, p1, p2, sy2, L, mu, Uv)
    a, b = batched.gamma_posterior(1.0, 1.0, N, s1, s2)
    F = batched.free_energy_regression(N, s1, s2, (a, b), (1.0, 1.0), mu, Sig, np.zeros(M), np.eye(M))
    meta = unisgp.UniSGPMeta(None, Z, None, None, np.zeros((M, M)), L, kern, Uv, 0, N)
    U = sum(unisgp.average_energy_pointmass(y[n], 0.0, X[n], mu, (a, b), th, meta) for n in range(N))
    assert abs(F - (U + batched.kl_mvn(mu, Sig, np.zeros(M), np.eye(M)) + batched.kl_gamma(a, b, 1.0, 1.0))) < 1e-9 * abs(F)
    assert np.isfinite(F)


def test_multisgp_in_message_is_the_negative_energy_up_to_its_constant():
    # `@rule MultiSGP(:in)` (MultiSGPnode.jl:162-211) returns exp(E[log p]) as a function of the input, `@average_energy` (:574-602) is
    # -E[log p] with the input marginal collapsed to a point: U(x) = D/2 ln 2pi - 1/2 ln|W| + 1/2 tr(W R_y) - log_backwardmess(x).
    # Ties the oracle's restatement of the :in rule to its energy rule (itself pinned by GPtest.jl's formulas).
    from oracle import multisgp, cubature as cub, kernels as ker
    rng = np.random.default_rng(3)
    M, D, d = 9, 2, 2
    Z = rng.normal(size=(M, d)) * 1.5
    theta = np.array([0.3, 0.7, 1.1])
    kern = lambda t: (ker.softplus(t[0]), ker.softplus(t[1:]), 0)
    var, ell, _ = kern(theta)
    Kinv = np.linalg.inv(ker.kuu(Z, var, ell, jitter=1e-10))
    meta = multisgp.MultiSGPMeta(cub.SRCUBATURE, Z, np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), Kinv, kern)
    W = np.array([[3.0, 0.4], [0.4, 2.0]])
    mu_v = rng.normal(size=D * M); C = rng.normal(size=(D * M, D * M)) * 0.2; Sigma_v = C @ C.T + 0.05 * np.eye(D * M)
    mu_y = rng.normal(size=D)
    f = multisgp.rule_in_logpdf(mu_y, mu_v, Sigma_v, W, theta, meta)
    const = 0.5 * D * np.log(2 * np.pi) - 0.5 * np.log(np.linalg.det(W)) + 0.5 * mu_y @ W @ mu_y
    for _ in range(5):
        x = rng.normal(size=d)
        U = multisgp.average_energy(mu_y, None, (x, 1e-30 * np.eye(d)), mu_v, Sigma_v, W, np.log(np.linalg.det(W)), theta, meta)
        assert abs(U - (const - f(x))) <= 1e-10 * max(abs(U), 1.0)
