"""Per-point restatement of the UniSGP node rules (oracle; test infrastructure only).

Follows GPnode/UniSGPnode.jl rule by rule, in the reference's own schedule: one message per data point,
folded left-to-right by ``prod`` into the running Gaussian, with the posterior factorised on the N-th fold.
Jitters and clamps are reproduced where (and only where) the reference has them.
"""
from dataclasses import dataclass, field
from typing import Any, Callable, Optional

import numpy as np
from scipy.linalg import solve_triangular
from scipy.special import digamma

from . import cubature as cub
from .kernels import kernel_matrix

LOG2PI = float(np.log(2.0 * np.pi))


@dataclass
class UniSGPMeta:
    """helper_functions/gp_helperfunction.jl:33-44 (same field names and order).
    ``kernel`` maps theta -> (variance, lengthscale, kind); ``method`` is None or (method_id, p)."""
    method: Any
    Xu: np.ndarray
    Psi0: np.ndarray
    Psi1_trans: np.ndarray
    Psi2: np.ndarray
    KuuL: np.ndarray
    kernel: Callable
    Uv: np.ndarray
    counter: int = 0
    N: int = 0


@dataclass
class BufferUniSGP:
    """GPnode/UniSGPnode.jl:56-59: (xi, Lambda) of a weighted-mean/precision Gaussian + the shared meta."""
    xi: np.ndarray
    Lam: np.ndarray
    meta: UniSGPMeta


def _col(meta, theta, x):
    var, ell, kind = meta.kernel(theta)
    Z = np.asarray(meta.Xu, dtype=np.float64)
    if Z.ndim == 1:
        Z = Z[:, None]
    return kernel_matrix(np.atleast_2d(np.asarray(x, dtype=np.float64).reshape(1, -1)), Z, var, ell, kind)[0], var


def _points(meta, m, v):
    mid, p = meta.method if isinstance(meta.method, tuple) else (meta.method, 21)
    return cub.sigma_points(mid, np.atleast_1d(m), np.atleast_2d(v), p)


def kernel_expectations(meta, theta, m, v):
    """approximate_kernel_expectation (UniSGPnode.jl:25-33): (Psi0, Psi1 [M], Psi2 [M x M]) under q(x)=N(m,v)."""
    pts, wts = _points(meta, m, v)
    psi0 = 0.0
    psi1 = 0.0
    psi2 = 0.0
    for pt, wt in zip(pts, wts):
        k, var = _col(meta, theta, pt)
        psi0 = psi0 + wt * var
        psi1 = psi1 + wt * k
        psi2 = psi2 + wt * np.outer(k, k)
    return psi0, psi1, psi2


# ---- :v rules -------------------------------------------------------------------------------------------
def rule_v_pointmass(mu_y, x, w, theta, meta):
    """UniSGPnode.jl:144-158 (regression) and :161-173 (classification, mu_y = E[f]).  The returned precision
    ALIASES meta.Psi2, as in the reference (:155,157)."""
    k, _ = _col(meta, theta, x)
    np.multiply(np.outer(k, k), w, out=meta.Psi2)          # mul!(meta.Psi2, k, k', w, 0)
    return BufferUniSGP(k * (mu_y * w), meta.Psi2, meta)


def rule_v_uncertain(mu_y, q_in, w, theta, meta):
    """UniSGPnode.jl:125-140: cubature expectations, + 1e-8 I on Psi2."""
    _, psi1, psi2 = kernel_expectations(meta, theta, q_in[0], q_in[1])
    psi2 = psi2 + 1e-8 * np.eye(psi2.shape[0])
    return BufferUniSGP(psi1 * (mu_y * w), psi2 * w, meta)


def prod_fold(left, right: BufferUniSGP):
    """UniSGPnode.jl:62-73.  ``left`` = (xi, Lambda) of the running Gaussian.  Returns the new (xi, Lambda);
    on the meta.N-th call refreshes meta.Uv = chol(Sigma_v + mu_v mu_v').U and resets the counter."""
    xi = left[0] + right.xi
    Lam = left[1] + right.Lam
    meta = right.meta
    meta.counter += 1
    if meta.counter == meta.N:
        mu_v, Sigma_v = mean_cov(xi, Lam)
        meta.Uv = np.linalg.cholesky(Sigma_v + np.outer(mu_v, mu_v)).T
        meta.counter = 0
    return xi, Lam


def mean_cov(xi, Lam):
    """ReactiveMP mean_cov of a weighted-mean/precision Gaussian: Sigma = cholinv(Lambda), mu = Sigma xi."""
    L = np.linalg.cholesky(0.5 * (Lam + Lam.T))
    Linv = solve_triangular(L, np.eye(L.shape[0]), lower=True)
    Sigma = Linv.T @ Linv
    return Sigma @ xi, Sigma


# ---- :w rules -------------------------------------------------------------------------------------------
def _I1_I2_point(mu_y, v_y, x, mu_v, theta, meta):
    k, var = _col(meta, theta, x)
    alpha = solve_triangular(meta.KuuL, k, lower=True)
    I1 = var - alpha @ alpha
    beta = meta.Uv @ k
    I2 = mu_y * mu_y + v_y - 2.0 * mu_y * (k @ mu_v) + beta @ beta
    return I1, I2


def rule_w_pointmass(mu_y, v_y, x, mu_v, theta, meta):
    """UniSGPnode.jl:196-216 (v_y = 0) and :219-238: returns (shape, rate) of the Gamma message."""
    I1, I2 = _I1_I2_point(mu_y, v_y, x, mu_v, theta, meta)
    return 1.5, 0.5 * (I1 + I2)


def _I1_I2_uncertain(mu_y, v_y, q_in, mu_v, theta, meta):
    psi0, psi1, psi2 = kernel_expectations(meta, theta, q_in[0], q_in[1])
    psi2 = psi2 + 1e-8 * np.eye(psi2.shape[0])
    A = solve_triangular(meta.KuuL, psi2, lower=True)
    A = solve_triangular(meta.KuuL.T, A, lower=False)
    I1 = np.clip(psi0 - np.trace(A), 1e-12, 1e12)
    I2 = np.clip(mu_y * mu_y + v_y - 2.0 * mu_y * (psi1 @ mu_v) + np.trace(meta.Uv.T @ meta.Uv @ psi2), 1e-12, 1e12)
    return I1, I2


def rule_w_uncertain(mu_y, v_y, q_in, mu_v, theta, meta):
    """UniSGPnode.jl:177-192 (clamps on I1, I2; + 1e-8 I on Psi2)."""
    I1, I2 = _I1_I2_uncertain(mu_y, v_y, q_in, mu_v, theta, meta)
    return 1.5, 0.5 * (I1 + I2)


# ---- :out rules -----------------------------------------------------------------------------------------
def rule_out_pointmass(x, mu_v, w_bar, theta, meta):
    """UniSGPnode.jl:96-104: NormalMeanPrecision(k' mu_v, w_bar) -> (mean, precision)."""
    k, _ = _col(meta, theta, x)
    return k @ mu_v, w_bar


def rule_out_uncertain(q_in, mu_v, w_bar, theta, meta):
    """UniSGPnode.jl:85-93."""
    _, psi1, _ = kernel_expectations(meta, theta, q_in[0], q_in[1])
    return psi1 @ mu_v, w_bar


# ---- average energy -------------------------------------------------------------------------------------
def gamma_stats(shape, rate):
    return shape / rate, float(digamma(shape) - np.log(rate))


def average_energy_pointmass(mu_y, v_y, x, mu_v, q_w, theta, meta):
    """UniSGPnode.jl:337-359 (regression, v_y = 0) and :363-387 (classification); q_w = (shape, rate)."""
    w_bar, E_logw = gamma_stats(*q_w)
    I1, I2 = _I1_I2_point(mu_y, v_y, x, mu_v, theta, meta)
    return 0.5 * (I1 * w_bar - E_logw + LOG2PI + I2 * w_bar)


def average_energy_uncertain(mu_y, v_y, q_in, mu_v, q_w, theta, meta):
    """UniSGPnode.jl:290-313."""
    w_bar, E_logw = gamma_stats(*q_w)
    I1, I2 = _I1_I2_uncertain(mu_y, v_y, q_in, mu_v, theta, meta)
    return 0.5 * (I1 * w_bar - E_logw + LOG2PI + I2 * w_bar)


def average_energy_pointmass_wpoint(mu_y, x, mu_v, Sigma_v, w_bar, theta, meta):
    """UniSGPnode.jl:411-436 (q_out, q_in, q_w all PointMass): recomputes chol(Sigma_v + mu mu').U itself."""
    k, var = _col(meta, theta, x)
    alpha = solve_triangular(meta.KuuL, k, lower=True)
    I1 = var - alpha @ alpha
    Lu = np.linalg.cholesky(Sigma_v + np.outer(mu_v, mu_v)).T
    beta = Lu @ k
    I2 = mu_y * mu_y - 2.0 * mu_y * (k @ mu_v) + beta @ beta
    return 0.5 * (I1 * w_bar - np.log(w_bar) + LOG2PI + I2 * w_bar)


def average_energy_gaussout_wpoint(mu_y, v_y, x, mu_v, Sigma_v, w_bar, theta, meta):
    """UniSGPnode.jl:438-458: ``.+ 1e-8`` on EVERY element of K_uu, Psi1 and Psi2, plain inv, clamps."""
    var, ell, kind = meta.kernel(theta)
    Z = np.asarray(meta.Xu, dtype=np.float64)
    Z = Z[:, None] if Z.ndim == 1 else Z
    Kuu_inv = np.linalg.inv(kernel_matrix(Z, Z, var, ell, kind) + 1e-8)
    k, _ = _col(meta, theta, x)
    psi1 = k + 1e-8
    psi2 = np.outer(k, k) + 1e-8
    I1 = np.clip(var - np.trace(Kuu_inv @ psi2), 1e-12, 1e12)
    I2 = np.clip(mu_y**2 + v_y - 2 * mu_y * (psi1 @ mu_v) + np.trace((Sigma_v + np.outer(mu_v, mu_v)) @ psi2), 1e-12, 1e12)
    return 0.5 * (I1 * w_bar - np.log(w_bar) + LOG2PI + I2 * w_bar)


def average_energy_uncertain_wpoint(mu_y, v_y, q_in, mu_v, Sigma_v, w_bar, theta, meta):
    """UniSGPnode.jl:390-409 (q_in Gaussian, q_w PointMass): cubature expectations, then ``.+ 1e-8`` on EVERY element of K_uu, Psi1 and
    Psi2, plain inv, clamps."""
    var, ell, kind = meta.kernel(theta)
    Z = np.asarray(meta.Xu, dtype=np.float64)
    Z = Z[:, None] if Z.ndim == 1 else Z
    Kuu_inv = np.linalg.inv(kernel_matrix(Z, Z, var, ell, kind) + 1e-8)
    psi0, psi1, psi2 = kernel_expectations(meta, theta, q_in[0], q_in[1])
    psi1 = psi1 + 1e-8
    psi2 = psi2 + 1e-8
    I1 = np.clip(psi0 - np.trace(Kuu_inv @ psi2), 1e-12, 1e12)
    I2 = np.clip(mu_y**2 + v_y - 2 * mu_y * (psi1 @ mu_v) + np.trace((Sigma_v + np.outer(mu_v, mu_v)) @ psi2), 1e-12, 1e12)
    return 0.5 * (I1 * w_bar - np.log(w_bar) + LOG2PI + I2 * w_bar)


# ---- the reference's sweep, as scheduled ----------------------------------------------------------------
def sweep_v_pointmass(X, ybar, w, theta, meta, prior_mean, prior_cov):
    """One ``infer(iterations=1)`` pass of experiments/regression_kin40k.ipynb:147-152,185-192: prior
    N(mu0, Sigma0) folded with the N per-point :v messages in data order.  Returns (mu_v, Sigma_v, xi, Lambda)."""
    X = np.asarray(X, dtype=np.float64)
    X = X[:, None] if X.ndim == 1 else X
    Lam0 = np.linalg.inv(prior_cov)
    state = (Lam0 @ prior_mean, Lam0.copy())
    meta.N = X.shape[0]
    meta.counter = 0
    for n in range(X.shape[0]):
        state = prod_fold(state, rule_v_pointmass(ybar[n], X[n], w, theta, meta))
    mu_v, Sigma_v = mean_cov(*state)
    return mu_v, Sigma_v, state[0], state[1]
