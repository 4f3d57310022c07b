"""Parity of the M x M factorisations, posterior update, :w terms and prediction against the oracle / LAPACK.
Solves are judged by scaled residual (backward error): two backward-stable solvers legitimately differ by cond * eps
in the forward error ('Conditioning vs the 1e-10 gate')."""
import numpy as np
import pytest
from scipy.linalg import solve_triangular

from oracle import batched, kernels

pytestmark = pytest.mark.gpu


def fro(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def ctx():
    from gaussianprocessnode_b200 import SGPContext
    c = SGPContext(0)
    yield c
    c.close()


@pytest.mark.parametrize("M,D", [(1, 1), (20, 1), (48, 2), (64, 2), (65, 3), (200, 8), (512, 8), (600, 8), (1024, 8)])
def test_kuu_factor_and_solve(ctx, M, D):
    rng = np.random.default_rng(M)
    Z = rng.normal(size=(M, D)) * (3.0 if D <= 2 else 1.0)
    ell = np.full(D, 0.8)
    jitter = 1e-8
    ctx.set_kernel(1.3, ell, D=D); ctx.set_inducing(Z)
    L = ctx.kuu_factor(jitter)
    K = kernels.kuu(Z, 1.3, ell, jitter=jitter)
    assert np.allclose(np.triu(L, 1), 0.0)
    assert fro(L @ L.T, K) < 1e-13                                   # factorisation residual
    if np.linalg.cond(K) < 1e6:
        assert fro(L, np.linalg.cholesky(K)) < 1e-10                 # factor itself when well conditioned
    B = rng.normal(size=(M, 3))
    Xs = ctx.kuu_solve(B)
    resid = np.linalg.norm(K @ Xs - B) / (np.linalg.norm(K) * np.linalg.norm(Xs) + np.linalg.norm(B))
    assert resid < 1e-13


def test_not_positive_definite_is_an_error(ctx):
    from gaussianprocessnode_b200 import SGPError
    Z = np.zeros((8, 1))                                             # identical inducing points, no jitter: singular
    ctx.set_kernel(1.0, np.array([1.0])); ctx.set_inducing(Z)
    with pytest.raises(SGPError) as e:
        ctx.kuu_factor(0.0)
    assert e.value.code == -3


@pytest.mark.parametrize("N,D,M,w", [(50, 1, 20, 25.0), (300, 2, 33, 10.0), (2000, 2, 64, 3.0), (3000, 3, 100, 30.0), (5000, 8, 256, 100.0),
                                     (10000, 8, 512, 1.0e4), (6000, 8, 600, 1.0e3), (4000, 8, 1000, 50.0), (3000, 6, 2100, 20.0)])
def test_posterior_and_w_terms(ctx, N, D, M, w):
    rng = np.random.default_rng(N + M)
    X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=N)
    Z = X[rng.choice(N, M, replace=False)] if D > 1 else np.linspace(-2.5, 2.5, M)[:, None]
    ell = np.full(D, 1.5 if D > 1 else 0.6); var = 1.2; jitter = 1e-6
    ctx.set_kernel(var, ell, D=D); ctx.set_inducing(Z); ctx.set_data(X, y)
    psi0, psi1, psi2, sy2 = ctx.sweep_psi()
    xi0 = np.zeros(M); Lam0 = np.eye(M) / 50.0                       # prior N(0, 50 I) (regression_kin40k.ipynb:200-201)
    mu, Sigma, Uv = ctx.posterior_v(xi0, Lam0, w)
    o_mu, o_Sig, o_Uv, o_Lam, o_xi = batched.posterior_v(xi0, Lam0, w, psi1, psi2)
    # backward errors: Lambda Sigma = I, Lambda mu = xi, Uv'Uv = Sigma + mu mu'
    nL = np.linalg.norm(o_Lam, 2)
    assert np.linalg.norm(o_Lam @ Sigma - np.eye(M)) / (nL * np.linalg.norm(Sigma, 2) * M) < 1e-13
    assert np.linalg.norm(o_Lam @ mu - o_xi) / (nL * np.linalg.norm(mu) + np.linalg.norm(o_xi)) < 1e-11
    cond = np.linalg.cond(o_Lam)
    # Uv comes from the factor of Sigma and p = U0 xi, not from the rounded Sigma and mu: its residual against THEM carries mu's own
    # conditioning-limited error (mu = Sigma xi, |Sigma| |xi| >> |mu|)
    assert fro(Uv.T @ Uv, Sigma + np.outer(mu, mu)) < max(1e-12, 20 * cond * 2.2e-16)
    assert np.allclose(np.tril(Uv, -1), 0.0)
    assert np.all(np.diag(Uv) > 0.0)
    Rv = o_Sig + np.outer(o_mu, o_mu)
    assert fro(Uv, o_Uv) < max(1e-10, 50 * np.linalg.cond(Rv) * 2.2e-16)   # the factor is unique: forward error bounded by conditioning
    assert fro(Sigma, o_Sig) < max(1e-10, 50 * cond * 2.2e-16)        # forward error bounded by conditioning
    assert fro(mu, o_mu) < max(1e-10, 50 * cond * 2.2e-16)
    # :w terms with the posterior as input
    L = ctx.kuu_factor(jitter)
    s1, s2 = ctx.w_terms(o_mu, o_Uv)
    Lo = np.linalg.cholesky(kernels.kuu(Z, var, ell, jitter=jitter))
    o1, o2 = batched.w_terms(psi0, psi1, psi2, sy2, Lo, o_mu, o_Uv)
    scale = abs(psi0) + abs(sy2)
    # tr(K_uu^-1 Psi2) inherits cond(K_uu) * eps from ANY backward-stable solver: tolerance is conditioning-aware
    condK = np.linalg.cond(kernels.kuu(Z, var, ell, jitter=jitter))
    tol1 = max(1e-10 * scale, 50 * condK * 2.2e-16 * abs(psi0 - o1))
    assert abs(s1 - o1) < tol1, (s1, o1, condK)
    assert abs(s2 - o2) < 1e-10 * max(abs(o2), scale), (s2, o2)
    # free energy of the regression model (<= 1e-8 relative, north_star)
    a, b = batched.gamma_posterior(1.0, 1.0, N, o1, o2)
    a2, b2 = batched.gamma_posterior(1.0, 1.0, N, s1, s2)
    F_o = batched.free_energy_regression(N, o1, o2, (a, b), (1.0, 1.0), o_mu, o_Sig, np.zeros(M), 50.0 * np.eye(M))
    F_g = batched.free_energy_regression(N, s1, s2, (a2, b2), (1.0, 1.0), mu, 0.5 * (Sigma + Sigma.T), np.zeros(M), 50.0 * np.eye(M))
    assert abs(F_g - F_o) <= 1e-8 * abs(F_o) + (a / b) * tol1, (F_g, F_o)   # + the conditioning-limited part of sumI1


@pytest.mark.parametrize("M", [48, 200, 512])
def test_posterior_without_uv_feeds_the_resident_consumers(ctx, M):
    # Uv == NULL skips the second Cholesky factorisation; the resident consumers use <R_v, Psi2> = <Sigma_v, Psi2> + mu_v' Psi2 mu_v
    rng = np.random.default_rng(M)
    N, D, w = 3000, 4, 40.0
    X = rng.normal(size=(N, D)); y = np.cos(X[:, 1]) + 0.1 * rng.normal(size=N)
    Z = X[rng.choice(N, M, replace=False)]
    ctx.set_kernel(0.9, np.full(D, 1.4)); ctx.set_inducing(Z); ctx.set_data(X, y)
    ctx.sweep_psi(fetch=False)
    ctx.kuu_factor(1e-6, fetch=False)
    ctx.prior_set_isotropic(50.0)
    mu, Sig, Uv = ctx.posterior_v_stream(w, carry=False, fetch=True)
    ref_w = ctx.w_terms(mu, Uv); ref_t = ctx.theta_objective(mu, Uv, w, 1e-6)
    mu2, Sig2, none = ctx.posterior_v_stream(w, carry=False, fetch=True, want_Uv=False)
    assert none is None and np.array_equal(mu, mu2) and np.array_equal(Sig, Sig2)
    got_w = ctx.w_terms(None, None); got_t = ctx.theta_objective(None, None, w, 1e-6)
    assert abs(got_w[0] - ref_w[0]) <= 1e-12 * abs(ref_w[0]) and abs(got_w[1] - ref_w[1]) <= 1e-10 * abs(ref_w[1])
    assert abs(got_t[0] - ref_t[0]) <= 1e-10 * abs(ref_t[0]) and abs(got_t[1] - ref_t[1]) <= 1e-9 * abs(ref_t[1])
    assert np.linalg.norm(got_t[2] - ref_t[2]) <= 1e-9 * np.linalg.norm(ref_t[2])


def test_theta_objective_leaves_the_resident_statistics_intact(ctx):
    # the theta step reuses (or re-creates) the statistics of the resident data: a following sgp_w_terms sees the same numbers
    rng = np.random.default_rng(77)
    N, D, M, w = 2000, 3, 96, 25.0
    X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]); yv = rng.uniform(0.0, 0.3, N)
    Z = X[:M].copy()
    ctx.set_kernel(1.1, np.full(D, 1.2)); ctx.set_inducing(Z); ctx.set_data(X, y, yv)
    p0, p1, p2, sy = ctx.sweep_psi()
    ctx.kuu_factor(1e-6, fetch=False); ctx.prior_set_isotropic(10.0)
    mu, Sig, Uv = ctx.posterior_v_stream(w, carry=False, fetch=True)
    before = ctx.w_terms(mu, Uv)
    ctx.theta_objective(mu, Uv, w, 1e-6)
    assert ctx.w_terms(mu, Uv) == before            # same bits: sum_y2 still carries Var[y], nothing was re-swept
    ctx.set_targets(y, yv)                          # invalidates the statistics: the theta step sweeps again, WITH the variances
    ctx.theta_objective(mu, Uv, w, 1e-6)
    after = ctx.w_terms(mu, Uv)
    assert after == before


def test_kin40k_golden_chain_on_gpu(ctx, kin40k):
    """K_*u mu_v over the 30000 test points with the reference's saved posterior reproduces the notebook's SMSE."""
    sp = kernels.softplus(kin40k["theta_raw"])
    Xu = kin40k["xtrain"][kin40k["xu_ids"]]
    ctx.set_kernel(sp[0], sp[1:]); ctx.set_inducing(Xu)
    pred = ctx.predict_mean(kin40k["xtest"], kin40k["mu_v"])
    assert abs(batched.smse(kin40k["ytest"], pred) - 0.08343114079545057) < 1e-12
    assert fro(pred, batched.predict_mean(kin40k["xtest"], Xu, sp[0], sp[1:], kin40k["mu_v"])) < 1e-12


def test_banana_golden_chain_on_gpu(ctx, banana):
    sp = kernels.softplus(banana["theta_raw"])
    x = banana["x"]; Xu = x[:4000][banana["xu_ids"]]
    ctx.set_kernel(sp[0], sp[1:]); ctx.set_inducing(Xu)
    m = ctx.predict_mean(x[4000:5300], banana["mu_v"])
    yt = (banana["label"][4000:5300] > 0).astype(float)
    assert int(np.sum(np.abs((m > 0).astype(float) - yt))) == 125
    # the notebook's own prediction path (classification_banana.ipynb:289-317): Probit(:out) of the :out message, decision at p >= 0.5
    from scipy.stats import norm
    w_bar = 3.7
    mf, var_f, prob = ctx.predict_probit(x[4000:5300], banana["mu_v"], w_bar)
    assert np.array_equal(mf, m) and var_f == 1.0 / w_bar
    assert np.max(np.abs(prob - norm.cdf(m / np.sqrt(1.0 + 1.0 / w_bar)))) < 1e-14
    assert int(np.sum(np.abs((prob >= 0.5).astype(float) - yt))) == 125


def test_streaming_prior_equals_one_full_sweep(ctx, kin40k):
    # experiments/regression_kin40k.ipynb:196-228: 20 mini-batches of 500, the posterior of batch b is the prior of batch b+1,
    # starting from N(0, 50 I).  For fixed theta that is the same posterior as ONE sweep over all 10000 points.
    from gaussianprocessnode_b200 import nodes as nd
    X, y = kin40k["xtrain"], kin40k["ytrain"]
    sp = kernels.softplus(kin40k["theta_raw"])
    M = 256
    Z = X[kin40k["xu_ids"][:M]]
    w = 1.0e4
    ctx.set_kernel(sp[0], sp[1:]); ctx.set_inducing(Z)
    ctx.prior_set_isotropic(50.0)
    xb, yb = nd.split2batch((X, y), 500)
    for b in range(len(xb)):
        ctx.set_data(xb[b], yb[b]); ctx.sweep_psi(fetch=False)
        ctx.posterior_v_stream(w, carry=True)                       # nothing crosses the bus
    mu_s, Sig_s, Uv_s = ctx.posterior_v_stream(0.0, carry=False, fetch=True)     # w = 0: posterior == resident prior
    ctx.set_data(X, y); ctx.sweep_psi(fetch=False)
    mu_f, Sig_f, Uv_f = ctx.posterior_v(np.zeros(M), np.eye(M) / 50.0, w)
    assert fro(mu_s, mu_f) < 1e-7 and fro(Sig_s, Sig_f) < 1e-7          # conditioning of Lambda (cond ~ 1e8: SURVEY.md 7)
    # the resident posterior feeds the :w terms and the theta step without host copies
    ctx.kuu_factor(1e-8, fetch=False)
    a = ctx.w_terms(None, None); b_ = ctx.w_terms(mu_s, Uv_s)
    assert abs(a[0] - b_[0]) <= 1e-12 * abs(b_[0]) and abs(a[1] - b_[1]) <= 1e-10 * abs(b_[1])
    f1 = ctx.theta_objective(None, None, w, 1e-8); f2 = ctx.theta_objective(mu_s, Uv_s, w, 1e-8)
    assert abs(f1[0] - f2[0]) <= 1e-10 * abs(f2[0]) and np.linalg.norm(f1[2] - f2[2]) <= 1e-9 * np.linalg.norm(f2[2])


_UV2_SCRIPT = """
import sys, numpy as np
from gaussianprocessnode_b200 import SGPContext
M = int(sys.argv[1]); rng = np.random.default_rng(7 + M)
X = rng.normal(size=(4000, 4)); y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=4000)
Z = X[rng.choice(4000, M, replace=False)]
c = SGPContext(0); c.set_kernel(1.1, np.full(4, 1.4), D=4); c.set_inducing(Z); c.set_data(X, y); c.sweep_psi()
mu, Sig, Uv = c.posterior_v(np.zeros(M), np.eye(M) / 50.0, 40.0)
np.savez(sys.argv[2], mu=mu, Sig=Sig, Uv=Uv); c.close()
"""


@pytest.mark.parametrize("M", [50, 130, 512])
def test_one_launch_posterior_agrees_with_the_two_factorisation_form(tmp_path, M):
    # default: reversed-order factorisation + closed-form rank-one update inside one launch; SGP_DENSE_UV2=1: Cholesky of Sigma + mu mu'
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for tag, val in (("one", "0"), ("two", "1")):
        f = str(tmp_path / f"{tag}.npz")
        env = dict(os.environ, SGP_DENSE_UV2=val, PYTHONPATH=root)
        subprocess.run([sys.executable, "-c", _UV2_SCRIPT, str(M), f], check=True, env=env, cwd=root, timeout=300)
        out[tag] = np.load(f)
    tol = max(1e-10, 50 * np.linalg.cond(out["two"]["Sig"]) * 2.2e-16)       # two backward-stable solvers: forward difference ~ cond * eps
    errs = [fro(out["one"][k], out["two"][k]) for k in ("Sig", "mu", "Uv")]
    assert max(errs) < tol, (errs, tol)
    assert np.allclose(np.tril(out["one"]["Uv"], -1), 0.0)


@pytest.mark.parametrize("M", [100, 512])
def test_kuu_job_in_the_shadow_of_the_posterior(ctx, M):
    # mini-batch schedule at a NEW theta: the stale K_uu is refactored inside the posterior's launch (two jobs, one cooperative kernel);
    # the consumers of K_uu^-1 must see exactly what an explicit sgp_kuu_factor gives
    rng = np.random.default_rng(3 * M)
    N, D, w, jit = 2000, 4, 30.0, 1e-6
    X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=N)
    Z = X[rng.choice(N, M, replace=False)]
    ctx.set_kernel(1.0, np.full(D, 1.2)); ctx.set_inducing(Z); ctx.set_data(X, y)
    ctx.kuu_factor(jit, fetch=False)                        # puts the jitter on record
    ell2 = np.full(D, 1.5)
    # explicit path
    ctx.set_kernel(0.8, ell2); ctx.sweep_psi(fetch=False); ctx.kuu_factor(jit, fetch=False); ctx.prior_set_isotropic(50.0)
    mu_a, Sig_a, Uv_a = ctx.posterior_v_stream(w, carry=False, fetch=True)
    w_a = ctx.w_terms(None, None); t_a = ctx.theta_objective(None, None, w, jit)
    # fused path: K_uu is stale when the posterior runs
    ctx.set_kernel(0.8, ell2); ctx.sweep_psi(fetch=False); ctx.prior_set_isotropic(50.0)
    mu_b, Sig_b, Uv_b = ctx.posterior_v_stream(w, carry=False, fetch=True)
    w_b = ctx.w_terms(None, None)                           # needs K_uu^-1: no explicit factorisation since set_kernel
    t_b = ctx.theta_objective(None, None, w, jit)
    assert np.array_equal(mu_a, mu_b) and np.array_equal(Sig_a, Sig_b) and np.array_equal(Uv_a, Uv_b)
    assert w_a == w_b
    assert t_a[0] == t_b[0] and t_a[1] == t_b[1] and np.array_equal(t_a[2], t_b[2])
    Xs = ctx.kuu_solve(np.eye(M)[:, :3])
    K = kernels.kuu(Z, 0.8, ell2, jitter=jit)
    assert np.linalg.norm(K @ Xs - np.eye(M)[:, :3]) / (np.linalg.norm(K) * np.linalg.norm(Xs)) < 1e-13


def test_plain_c_client_against_the_oracle(tmp_path):
    # examples/c_client.c: sweep + N-th prod + :w terms through the C ABI from plain C (no Python, no torch in the process)
    import subprocess
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("_abi_helpers", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_abi.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    build_c_client = mod.build_c_client
    N, M, D = 300, 40, 3
    r = subprocess.run([build_c_client(tmp_path), str(N), str(M), str(D)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = {ln.split()[0]: np.array([float(v) for v in ln.split()[1:]]) for ln in r.stdout.splitlines()[2:]}

    state = [12345]

    def lcg():
        state[0] = (state[0] * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        return ((state[0] >> 11) & ((1 << 53) - 1)) / float(1 << 52) - 1.0
    X = np.array([2.0 * lcg() for _ in range(N * D)]).reshape(N, D)
    y = np.array([np.sin(X[n, 0]) + 0.05 * lcg() for n in range(N)])
    Z = np.array([2.0 * lcg() for _ in range(M * D)]).reshape(M, D)
    ell = 1.0 + 0.1 * np.arange(D); var = 1.3; w = 25.0; jit = 1e-6
    psi0, psi1, psi2, sy2 = batched.psi_stats_point(X, y, Z, var, ell)
    assert abs(out["psi0"][0] - psi0) <= 1e-13 * abs(psi0) and abs(out["sum_y2"][0] - sy2) <= 1e-13 * abs(sy2)
    assert fro(out["psi1"], psi1) < 1e-13 and fro(out["psi2_diag"], np.diag(psi2)) < 1e-13
    o_mu, o_Sig, o_Uv, o_Lam, _ = batched.posterior_v(np.zeros(M), np.eye(M) / 50.0, w, psi1, psi2)
    cond = np.linalg.cond(o_Lam)
    assert fro(out["mu"], o_mu) < max(1e-10, 50 * cond * 2.2e-16)
    assert fro(out["Uv_diag"], np.diag(o_Uv)) < max(1e-10, 50 * np.linalg.cond(o_Sig + np.outer(o_mu, o_mu)) * 2.2e-16)
    Lo = np.linalg.cholesky(kernels.kuu(Z, var, ell, jitter=jit))
    o1, o2 = batched.w_terms(psi0, psi1, psi2, sy2, Lo, o_mu, o_Uv)
    scale = abs(psi0) + abs(sy2)
    condK = np.linalg.cond(kernels.kuu(Z, var, ell, jitter=jit))
    assert abs(out["sumI1"][0] - o1) < max(1e-10 * scale, 50 * condK * 2.2e-16 * abs(psi0 - o1))
    assert abs(out["sumI2"][0] - o2) < 1e-9 * max(abs(o2), scale)
