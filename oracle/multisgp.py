"""Per-point restatement of the MultiSGP node rules (oracle; test infrastructure only).

Follows GPnode/MultiSGPnode.jl (live, non-commented rules only): vector output of dimension D with Gaussian
(uncertain) input q(x)=N(m,S); kernel expectations by a sigma-point rule (MultiSGPnode.jl:11-35); q(v) over the
D*M stacked transformed inducing values, output-major blocks of M.  No jitter on Psi2.
"""
from dataclasses import dataclass
from typing import Any, Callable

import numpy as np

from . import cubature as cub
from .kernels import kernel_matrix

LOG2PI = float(np.log(2.0 * np.pi))


@dataclass
class MultiSGPMeta:
    """helper_functions/gp_helperfunction.jl:55-64 (GPCache omitted: scratch only)."""
    method: Any
    Xu: np.ndarray
    Psi0: np.ndarray
    Psi1_trans: np.ndarray
    Psi2: np.ndarray
    Kuu_inverse: np.ndarray
    kernel: Callable


def kernel_expectations(meta, theta, m, P):
    """approximate_kernel_expectation! (MultiSGPnode.jl:15-24) for Psi0, Psi1_trans (M), Psi2 (M x M)."""
    mid, p = meta.method if isinstance(meta.method, tuple) else (meta.method, 21)
    pts, wts = cub.sigma_points(mid, m, P, p)
    var, ell, kind = meta.kernel(theta)
    Z = np.asarray(meta.Xu, dtype=np.float64)
    psi0 = 0.0; psi1 = np.zeros(Z.shape[0]); psi2 = np.zeros((Z.shape[0], Z.shape[0]))
    for pt, wt in zip(pts, wts):
        k = kernel_matrix(pt[None, :], Z, var, ell, kind)[0]
        psi0 += wt * var
        psi1 += wt * k                      # axpy!(weight, g(point), gbar)
        psi2 += wt * np.outer(k, k)
    return psi0, psi1, psi2


def create_blockmatrix(A, d, M):
    """helper_functions/gp_helperfunction.jl:133-135."""
    return [[A[i * M:(i + 1) * M, j * M:(j + 1) * M] for j in range(d)] for i in range(d)]


def sum_diagonal_M(V, M):
    """helper_functions/derivative_helper.jl:119-122."""
    return sum(V[M * i:M * (i + 1), i] for i in range(V.shape[1]))


def rule_v(mu_y, q_in, W, theta, meta):
    """MultiSGPnode.jl:290-308 / :310-328 -> (xi [D*M], Lambda [D*M x D*M])."""
    _, psi1, psi2 = kernel_expectations(meta, theta, *q_in)
    meta.Psi1_trans[:, 0] = psi1; meta.Psi2[:] = psi2
    r = np.asarray(mu_y) @ W                                  # mu_y' W  (1 x D)
    xi = np.concatenate([psi1 * r[d] for d in range(r.size)])  # vcat(Psi1_trans .* (mu_y' W)...)
    return xi, np.kron(W, psi2)


def rule_out(q_in, mu_v, W, theta, meta):
    """MultiSGPnode.jl:90-120 -> (mean [D], precision W)."""
    _, psi1, _ = kernel_expectations(meta, theta, *q_in)
    M = psi1.size
    D = W.shape[0]
    return np.array([psi1 @ mu_v[d * M:(d + 1) * M] for d in range(D)]), W


def rule_w(mu_y, Sigma_y, q_in, mu_v, Sigma_v, theta, meta):
    """MultiSGPnode.jl:367-405 (Sigma_y given) / :407-444 (Sigma_y = None) -> (nu = D+2, inverse scale Psi_4)."""
    mu_y = np.asarray(mu_y, dtype=np.float64)
    D = mu_y.size
    psi0, psi1, psi2 = kernel_expectations(meta, theta, *q_in)
    M = psi1.size
    Rv = Sigma_v + np.outer(mu_v, mu_v)
    blk = create_blockmatrix(Rv, D, M)
    I1s = psi0 - np.trace(meta.Kuu_inverse @ psi2)
    E = np.array([psi1 @ mu_v[d * M:(d + 1) * M] for d in range(D)])
    Psi4 = np.array([[np.sum(blk[i][j] * psi2.T) for j in range(D)] for i in range(D)])
    tmp = np.outer(mu_y, E)
    tmp = tmp + tmp.T
    Psi4 = Psi4 + np.outer(mu_y, mu_y) + (0.0 if Sigma_y is None else Sigma_y)
    Psi4 = Psi4 - tmp + I1s * np.eye(D)
    return D + 2, Psi4


def average_energy(mu_y, Sigma_y, q_in, mu_v, Sigma_v, W_bar, E_logW, theta, meta):
    """MultiSGPnode.jl:544-571 (Wishart q_w, Gaussian q_out), :574-602 (PointMass), :604-631."""
    mu_y = np.asarray(mu_y, dtype=np.float64)
    D = mu_y.size
    psi0, psi1, psi2 = kernel_expectations(meta, theta, *q_in)
    M = psi1.size
    Rv = Sigma_v + np.outer(mu_v, mu_v)
    V = np.outer(mu_v, mu_y) @ W_bar
    sumdiagV = sum_diagonal_M(V, M)
    blk = create_blockmatrix(Rv, D, M)
    sumRvblk_W = sum(blk[i][j] * W_bar[i, j] for i in range(D) for j in range(D))
    Ry = np.outer(mu_y, mu_y) + (0.0 if Sigma_y is None else Sigma_y)
    return (0.5 * D * LOG2PI - 0.5 * E_logW + 0.5 * np.trace(W_bar @ Ry)
            + 0.5 * np.trace(W_bar) * (psi0 - np.sum(meta.Kuu_inverse * psi2))
            - np.sum(sumdiagV * psi1) + 0.5 * np.sum(psi2 * sumRvblk_W))


def rule_in_logpdf(mu_y, mu_v, Sigma_v, W, theta, meta):
    """MultiSGPnode.jl:162-185 (q_out Gaussian) / :187-211 (PointMass): the closure log_backwardmess(x), written exactly as the
    reference has it (Psi0, Psi1_trans, Psi2 closures; sumdiagV; sumRvblk_W)."""
    mu_y = np.asarray(mu_y, dtype=np.float64)
    D = mu_y.size
    var, ell, kind = meta.kernel(theta)
    Z = np.asarray(meta.Xu, dtype=np.float64)
    M = Z.shape[0]
    Rv = Sigma_v + np.outer(mu_v, mu_v)
    V = np.outer(mu_v, mu_y) @ W
    sumdiagV = sum_diagonal_M(V, M)
    blk = create_blockmatrix(Rv, D, M)
    sumRvblk_W = sum(blk[i][j] * W[i, j] for i in range(D) for j in range(D))

    def log_backwardmess(x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        psi0 = kernel_matrix(x[None, :], x[None, :], var, ell, kind)[0, 0]
        psi1 = kernel_matrix(x[None, :], Z, var, ell, kind)[0]
        psi2 = np.outer(psi1, psi1)
        return (-0.5 * np.trace(W) * (psi0 - np.sum(meta.Kuu_inverse * psi2)) + np.sum(sumdiagV * psi1) - 0.5 * np.sum(psi2 * sumRvblk_W))
    return log_backwardmess


def prod_gaussian_logpdf(m, P, logpdf):
    """`prod(::GenericProd, left::MvGaussian, right::ContinuousMultivariateLogPdf)` (MultiSGPnode.jl:38-45): moment matching with
    the spherical-radial cubature points of `left`; ReactiveMP's approximate_meancov restated from its definition (unvendored:
    parity unpinned):  Z = sum w g(x),  mean = sum w g(x) x / Z,  cov = sum w g(x) (x - mean)(x - mean)' / Z,  g = exp(logpdf).
    Returns `left` unchanged when the mean is NaN."""
    pts, wts = cub.sigma_points(cub.SRCUBATURE, m, P)
    g = np.array([np.exp(logpdf(pt)) for pt in pts])
    Zn = np.sum(wts * g)
    mean = (wts * g) @ pts / Zn
    if np.isnan(mean[0]):
        return np.asarray(m, dtype=np.float64), np.asarray(P, dtype=np.float64)
    dlt = pts - mean
    cov = (dlt * (wts * g)[:, None]).T @ dlt / Zn
    return mean, cov


def rule_in_laplace(mu_y, x0, mu_v, Sigma_v, W, theta, meta, iters=200):
    """MultiSGPnode.jl:213-236: mode m_z of the backward message from x0 = mean(q_in) and W_z = Hessian of the negative log
    message there -> (xi = W_z m_z, W_z).  The reference runs Optim's LBFGS for 20 iterations with a ForwardDiff gradient and
    takes Zygote.hessian; Optim is not vendored (parity unpinned), so the oracle converges scipy's L-BFGS-B instead and
    differentiates numerically: the two agree wherever the reference's 20 iterations have converged."""
    from scipy.optimize import minimize
    f = rule_in_logpdf(mu_y, mu_v, Sigma_v, W, theta, meta)
    neg = lambda x: -f(x)
    res = minimize(neg, np.asarray(x0, dtype=np.float64), method="L-BFGS-B", options=dict(maxiter=iters, ftol=1e-15, gtol=1e-12))
    mz = res.x
    d = mz.size
    h = 1e-4
    H = np.zeros((d, d))
    for i in range(d):
        for j in range(d):
            ei = np.eye(d)[i] * h; ej = np.eye(d)[j] * h
            H[i, j] = (neg(mz + ei + ej) - neg(mz + ei - ej) - neg(mz - ei + ej) + neg(mz - ei - ej)) / (4 * h * h)
    return H @ mz, H, mz
