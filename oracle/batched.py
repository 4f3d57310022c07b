"""Batched identities the per-point rules collapse to (oracle; test infrastructure only).

These are what the CUDA path computes in one sweep.  ``tests/test_oracle_rules.py`` proves them equal to the
per-point schedule of ``oracle.unisgp`` / ``oracle.multisgp`` to ~1e-13 -- that equality IS the parity argument
for replacing N rule invocations by one kernel launch.
"""
import numpy as np
from scipy.linalg import solve_triangular
from scipy.special import digamma, multigammaln

from . import cubature as cub
from .kernels import kernel_matrix, SE

LOG2PI = float(np.log(2.0 * np.pi))


def _2d(X):
    X = np.asarray(X, dtype=np.float64)
    return X[:, None] if X.ndim == 1 else X


def psi_stats_point(X, ybar, Z, variance, ell, kind=SE, weights=None, yvar=None):
    """Psi0 = sum_n w_n k_nn, Psi1 = K_uf (w*ybar), Psi2 = K_uf diag(w) K_uf', sum_y2 = sum_n w_n (ybar_n^2 + yvar_n).
    (rows a2, a5 of SURVEY.md section 8; UniSGPnode.jl:144-173, 196-238 summed over n)."""
    X = _2d(X); Z = _2d(Z)
    K = kernel_matrix(X, Z, variance, ell, kind)            # N x M
    N = X.shape[0]
    w = np.ones(N) if weights is None else np.asarray(weights, dtype=np.float64)
    ybar = np.zeros(N) if ybar is None else np.asarray(ybar, dtype=np.float64)
    yv = np.zeros(N) if yvar is None else np.asarray(yvar, dtype=np.float64)
    psi0 = float(variance * np.sum(w))
    psi1 = K.T @ (w * ybar)
    psi2 = (K * w[:, None]).T @ K
    psi2 = 0.5 * (psi2 + psi2.T)
    sum_y2 = float(np.sum(w * (ybar * ybar + yv)))
    return psi0, psi1, psi2, sum_y2


def sigma_point_cloud(method, mean, cov, p=21):
    """All S sigma points of all N inputs: (points [N*S x d], weights [N*S], S).  cov: N x d x d (or N for d=1)."""
    mean = _2d(mean)
    N, d = mean.shape
    cov = np.asarray(cov, dtype=np.float64).reshape(N, d, d)
    P, W = [], []
    for n in range(N):
        pts, wts = cub.sigma_points(method, mean[n], cov[n], p)
        P.append(pts); W.append(wts)
    S = P[0].shape[0]
    return np.concatenate(P, 0), np.concatenate(W, 0), S


def psi_stats_uncertain(method, mean, cov, Z, variance, ell, kind=SE, p=21, YW=None, ybar=None):
    """Cubature Psi statistics summed over points (rows a3, a8, a11): Psi0 = sum_n sum_s om_s k(x_s,x_s),
    Psi1 = sum_n Psi1_n (x) r_n  (M x D_out with r_n = (YW)[n,:]; or M with r_n = ybar_n; or plain sum),
    Psi2 = sum_n sum_s om_s k_s k_s'.  Also returns the per-point Psi1_n (N x M)."""
    Z = _2d(Z)
    pts, wts, S = sigma_point_cloud(method, mean, cov, p)
    N = pts.shape[0] // S
    K = kernel_matrix(pts, Z, variance, ell, kind)          # (N*S) x M
    psi0 = float(variance * np.sum(wts))
    psi1_n = (K * wts[:, None]).reshape(N, S, -1).sum(1)    # N x M
    if YW is not None:
        psi1 = psi1_n.T @ np.asarray(YW, dtype=np.float64)  # M x D_out
    elif ybar is not None:
        psi1 = psi1_n.T @ np.asarray(ybar, dtype=np.float64)
    else:
        psi1 = psi1_n.sum(0)
    psi2 = (K * wts[:, None]).T @ K
    return psi0, psi1, 0.5 * (psi2 + psi2.T), psi1_n


def psi_stats_closed_form_se(mean, cov, Z, variance, ell, YW=None, ybar=None):
    """Closed-form SE-ARD Psi statistics under q(x_n)=N(m_n,S_n) -- an EXTENSION, not in
    the reference: its parity target is high-order Gauss-Hermite cubature, never the reference's srcubature."""
    mean = _2d(mean); Z = _2d(Z)
    N, d = mean.shape
    M = Z.shape[0]
    cov = np.asarray(cov, dtype=np.float64).reshape(N, d, d)
    lam = np.broadcast_to(np.asarray(ell, dtype=np.float64), (d,)) ** 2
    Lam = np.diag(lam)
    dz = (Z[:, None, :] - Z[None, :, :])
    Ezz = np.exp(-0.25 * np.einsum("abd,d,abd->ab", dz, 1.0 / lam, dz))
    zbar = 0.5 * (Z[:, None, :] + Z[None, :, :])
    psi1_n = np.empty((N, M)); psi2 = np.zeros((M, M))
    for n in range(N):
        S = cov[n]
        c1 = variance / np.sqrt(np.linalg.det(np.eye(d) + S / lam[:, None]))
        B1 = np.linalg.inv(Lam + S)
        dm = mean[n][None, :] - Z
        psi1_n[n] = c1 * np.exp(-0.5 * np.einsum("ad,de,ae->a", dm, B1, dm))
        c2 = variance**2 / np.sqrt(np.linalg.det(np.eye(d) + 2.0 * S / lam[:, None]))
        B2 = np.linalg.inv(Lam + 2.0 * S)
        dmz = mean[n][None, None, :] - zbar
        psi2 += c2 * Ezz * np.exp(-np.einsum("abd,de,abe->ab", dmz, B2, dmz))
    psi0 = float(variance * N)
    if YW is not None:
        psi1 = psi1_n.T @ np.asarray(YW, dtype=np.float64)
    elif ybar is not None:
        psi1 = psi1_n.T @ np.asarray(ybar, dtype=np.float64)
    else:
        psi1 = psi1_n.sum(0)
    return psi0, psi1, 0.5 * (psi2 + psi2.T), psi1_n


def cholinv(A):
    L = np.linalg.cholesky(0.5 * (A + A.T))
    Li = solve_triangular(L, np.eye(L.shape[0]), lower=True)
    return Li.T @ Li


def posterior_v(xi0, Lam0, w, psi1, psi2):
    """Fold of the prior with all N :v messages + the N-th-``prod`` flush (UniSGPnode.jl:62-73):
    Lambda = Lambda0 + w Psi2, xi = xi0 + w Psi1, Sigma_v = cholinv(Lambda), mu_v = Sigma_v xi,
    U_v = chol(Sigma_v + mu_v mu_v').U.  Returns (mu_v, Sigma_v, U_v, Lambda, xi)."""
    Lam = np.asarray(Lam0, dtype=np.float64) + w * psi2
    xi = np.asarray(xi0, dtype=np.float64) + w * psi1
    Sigma = cholinv(Lam)
    mu = Sigma @ xi
    Uv = np.linalg.cholesky(Sigma + np.outer(mu, mu)).T
    return mu, Sigma, Uv, Lam, xi


def w_terms(psi0, psi1, psi2, sum_y2, KuuL, mu_v, Uv):
    """sum_n I1_n = Psi0 - tr(K_uu^-1 Psi2);  sum_n I2_n = sum(ybar^2+v) - 2 mu_v'Psi1 + <U_v'U_v, Psi2>
    (UniSGPnode.jl:196-238 summed; trsv -> trsm on Psi2, trmv -> Frobenius product)."""
    A = solve_triangular(KuuL, psi2, lower=True)
    sumI1 = psi0 - float(np.sum(solve_triangular(KuuL, A.T, lower=True).diagonal()))
    B = Uv @ psi2
    sumI2 = sum_y2 - 2.0 * float(mu_v @ psi1) + float(np.sum(B * Uv))
    return sumI1, sumI2


def gamma_posterior(a0, b0, N, sumI1, sumI2):
    """prior Gamma(a0,b0) x N messages Gamma(1.5, r_n)."""
    return a0 + 0.5 * N, b0 + 0.5 * (sumI1 + sumI2)


def energy_sum(N, sumI1, sumI2, w_bar, E_logw):
    """sum_n U_n of UniSGPnode.jl:337-387: 0.5 w (sumI1+sumI2) + N/2 (ln 2pi - E ln w)."""
    return 0.5 * w_bar * (sumI1 + sumI2) + 0.5 * N * (LOG2PI - E_logw)


def kl_mvn(mu_q, Sigma_q, mu_p, Sigma_p):
    M = mu_q.size
    Lp = np.linalg.cholesky(Sigma_p); Lq = np.linalg.cholesky(Sigma_q)
    A = solve_triangular(Lp, Lq, lower=True)
    d = solve_triangular(Lp, mu_q - mu_p, lower=True)
    return 0.5 * (np.sum(A * A) + d @ d - M + 2.0 * (np.sum(np.log(np.diag(Lp))) - np.sum(np.log(np.diag(Lq)))))


def kl_gamma(a, b, a0, b0):
    from scipy.special import gammaln
    return (a - a0) * digamma(a) - gammaln(a) + gammaln(a0) + a0 * (np.log(b) - np.log(b0)) + a * (b0 - b) / b


def free_energy_regression(N, sumI1, sumI2, q_w, prior_w, mu_v, Sigma_v, mu0, Sigma0):
    """Bethe free energy of the regression model for fixed theta:
    sum_n U_n + KL(q(v)||p(v)) + KL(q(w)||p(w))."""
    a, b = q_w
    w_bar = a / b
    E_logw = float(digamma(a) - np.log(b))
    return energy_sum(N, sumI1, sumI2, w_bar, E_logw) + kl_mvn(mu_v, Sigma_v, mu0, Sigma0) + kl_gamma(a, b, *prior_w)


def predict_mean(Xt, Z, variance, ell, mu_v, kind=SE):
    """:out rule over a test set (UniSGPnode.jl:96-104; regression_kin40k.ipynb:289-304): K_*u mu_v."""
    return kernel_matrix(_2d(Xt), _2d(Z), variance, ell, kind) @ mu_v


def smse(y_true, y_pred):
    """helper_functions/gp_helperfunction.jl:145-149 (var with Julia's n-1 normalisation)."""
    y_true = np.asarray(y_true, dtype=np.float64); y_pred = np.asarray(y_pred, dtype=np.float64)
    mse = np.linalg.norm(y_true - y_pred) ** 2 / y_true.size
    return mse / np.var(y_true, ddof=1)


def probit_moments(m, v, y01):
    """Analytic q(f_n) for the Probit likelihood with cavity N(m, v) (ReactiveMP's Probit
    node is unvendored -> parity unpinned; the banana chain pins only the final decision sign)."""
    from scipy.special import log_ndtr
    s = 2.0 * np.asarray(y01, dtype=np.float64) - 1.0
    z = s * m / np.sqrt(1.0 + v)
    r = np.exp(-0.5 * z * z - 0.5 * np.log(2 * np.pi) - log_ndtr(z))
    Ef = m + s * v * r / np.sqrt(1.0 + v)
    Vf = v - v * v * r * (z + r) / (1.0 + v)
    return Ef, Vf


# ---- MultiSGP batched forms (rows a11-a14) ----------------------------------------------------------------
def multi_v_message(W_bar, psi1_mat, psi2):
    """sum_n of MultiSGPnode.jl:290-328: Lambda = kron(W, sum_n Psi2_n), xi = vec(Psi1bar (Y W)) with
    psi1_mat = sum_n Psi1_n (x) (W' mu_y_n) already contracted (M x D, column d = output d)."""
    return psi1_mat.T.reshape(-1).copy(), np.kron(W_bar, psi2)   # output-major blocks of M


def wishart_stats(nu, S):
    D = S.shape[0]
    sign, logdet = np.linalg.slogdet(S)
    E_logdet = float(np.sum(digamma(0.5 * (nu - np.arange(D)))) + D * np.log(2.0) + logdet)
    return nu * S, E_logdet
