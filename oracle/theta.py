"""The theta step's collapsed objective (oracle; test infrastructure only).

Restates helper_functions/derivative_helper.jl:23-39 (`neg_log_backwardmess_fast`) point by point, exactly as written, plus
the analytic gradient with respect to (variance, lengthscale) that the reference obtains with ForwardDiff
(derivative_helper.jl:55-67) -- forward-mode AD evaluates the exact derivative, so the analytic form is the parity target;
tests/test_oracle_theta.py pins it against central differences of the restated objective."""
import numpy as np
from scipy.linalg import solve_triangular

from .kernels import SE, MATERN32, MATERN52, kernel_matrix, kuu


def neg_log_backwardmess_fast(variance, ell, y_data, x_data, v, Uv, w, Xu, kind=SE, jitter=0.0):
    """derivative_helper.jl:23-39 with kernel(theta) -> (variance, ell, kind) already evaluated by the caller."""
    X = np.asarray(x_data, dtype=np.float64); X = X[:, None] if X.ndim == 1 else X
    Z = np.asarray(Xu, dtype=np.float64); Z = Z[:, None] if Z.ndim == 1 else Z
    Lu = np.linalg.cholesky(kuu(Z, variance, ell, kind, jitter))
    Kux = kernel_matrix(X, Z, variance, ell, kind).T          # M x N
    llh = 0.0
    for i in range(X.shape[0]):
        k = Kux[:, i]
        alpha = solve_triangular(Lu, k, lower=True)
        beta = Uv @ k
        llh += -0.5 * w * variance + 0.5 * w * (alpha @ alpha) - 0.5 * w * (beta @ beta) + w * y_data[i] * (v @ k)
    return -llh


def _h(kind, variance, r2):
    """d k / d ell_d = h(r) (x_d - z_d)^2 / ell_d^3."""
    if kind == SE:
        return variance * np.exp(-0.5 * r2)
    if kind == MATERN32:
        return variance * 3.0 * np.exp(-np.sqrt(3.0 * r2))
    s = np.sqrt(5.0 * r2)
    return variance * (5.0 / 3.0) * (1.0 + s) * np.exp(-s)


def objective_and_gradient(variance, ell, y_data, x_data, v, Uv, w, Xu, kind=SE, jitter=0.0):
    """(F, dF/dvariance, dF/dell[D]) in the batched form of DESIGN.md section 4.5."""
    X = np.asarray(x_data, dtype=np.float64); X = X[:, None] if X.ndim == 1 else X
    Z = np.asarray(Xu, dtype=np.float64); Z = Z[:, None] if Z.ndim == 1 else Z
    ell = np.broadcast_to(np.asarray(ell, dtype=np.float64), (Z.shape[1],))
    y = np.asarray(y_data, dtype=np.float64)
    N, M = X.shape[0], Z.shape[0]
    K = kernel_matrix(X, Z, variance, ell, kind)              # N x M
    Kuu = kuu(Z, variance, ell, kind, jitter)
    Kinv = np.linalg.inv(Kuu)
    Rv = Uv.T @ Uv
    psi0 = variance * N; psi1 = K.T @ y; psi2 = K.T @ K
    F = 0.5 * w * (psi0 - np.sum(Kinv * psi2) + np.sum(Rv * psi2)) - w * (v @ psi1)
    A = Rv - Kinv
    B = Kinv @ psi2 @ Kinv
    dvar = (0.5 * w * psi0 - 0.5 * w * np.sum(Kinv * psi2) + w * np.sum(Rv * psi2) - w * (v @ psi1) - 0.5 * w * jitter * np.trace(B)) / variance
    Cx = (X[:, None, :] - Z[None, :, :]) ** 2                 # N x M x D
    r2 = np.sum(Cx / ell ** 2, axis=2)
    Hx = _h(kind, variance, r2)
    G = K @ A                                                 # row n = (A k_n)'
    fac = Hx * (w * G - w * y[:, None] * v[None, :])
    dell = np.einsum("nm,nmd->d", fac, Cx)
    Cz = (Z[:, None, :] - Z[None, :, :]) ** 2
    r2z = np.sum(Cz / ell ** 2, axis=2)
    dell = dell + 0.5 * w * np.einsum("ab,abd->d", B * _h(kind, variance, r2z), Cz)
    return float(F), float(dvar), dell / ell ** 3
