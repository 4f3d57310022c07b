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
