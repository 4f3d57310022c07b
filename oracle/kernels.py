"""Covariance functions of the path (oracle; test infrastructure only).

Reference: every rule builds ``kernel(theta)`` and calls KernelFunctions' ``kernelmatrix(!)``
(GPnode/UniSGPnode.jl:153,169,206,229,350,377,424; GPnode/MultiSGPnode.jl:102,118,302-303).  The kernel is
``sigma2 * with_lengthscale(SEKernel(), ell)`` (experiments/regression_kin40k.ipynb:108, GPtest.jl:21), i.e.
``k(x,z) = sigma2 * exp(-0.5 * sum_d ((x_d - z_d)/ell_d)^2)`` -- KernelFunctions.jl (unvendored, unpinned); the
convention is pinned numerically by the kin40k golden chain.  Matern kernels are only ever *imported* by the
reference (regression_kin40k.ipynb:28); they follow KernelFunctions' convention r = ||(x-z)/ell||.
"""
import numpy as np

SE, MATERN32, MATERN52 = 0, 1, 2


def softplus(t):
    t = np.asarray(t, dtype=np.float64)
    return np.logaddexp(0.0, t)


def invsoftplus(s):
    s = np.asarray(s, dtype=np.float64)
    return s + np.log(-np.expm1(-s))


def _as2d(X, D=None):
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None] if (D is None or D == 1) else X[None, :]
    return X


def sqdist(X, Z, ell):
    """Scaled squared distances, direct-difference form (the numerically safest one). X: N x D, Z: M x D."""
    X = _as2d(X); Z = _as2d(Z)
    ell = np.broadcast_to(np.asarray(ell, dtype=np.float64), (X.shape[1],))
    d = (X[:, None, :] - Z[None, :, :]) / ell
    return np.einsum("nmd,nmd->nm", d, d)


def kernel_matrix(X, Z, variance, ell, kind=SE):
    """K[n, m] = k(x_n, z_m).  (N x M; the reference's Psi1_trans column for point n is K[n, :].)"""
    r2 = sqdist(X, Z, ell)
    if kind == SE:
        return variance * np.exp(-0.5 * r2)
    r = np.sqrt(r2)
    if kind == MATERN32:
        s = np.sqrt(3.0) * r
        return variance * (1.0 + s) * np.exp(-s)
    if kind == MATERN52:
        s = np.sqrt(5.0) * r
        return variance * (1.0 + s + s * s / 3.0) * np.exp(-s)
    raise ValueError("unknown kernel kind")


def kernel_diag(X, variance, ell=None, kind=SE):
    """k(x_n, x_n) (kernelmatrix_diag, helper_functions/derivative_helper.jl:26) = variance for all three."""
    X = _as2d(X)
    return np.full(X.shape[0], float(variance))


def kuu(Z, variance, ell, kind=SE, jitter=0.0):
    """K_uu (+ jitter*I): regression_kin40k.ipynb:183 (no jitter), classification_banana.ipynb:163 (1e-8),
    Pendulum_Wishart_2d.ipynb:2542 (1e-12)."""
    K = kernel_matrix(Z, Z, variance, ell, kind)
    K = 0.5 * (K + K.T)
    if jitter:
        K = K + jitter * np.eye(K.shape[0])
    return K
