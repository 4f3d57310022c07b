"""Host-side mirror of the reference's node interface for the hot path (GPnode/UniSGPnode.jl, GPnode/MultiSGPnode.jl,
helper_functions/gp_helperfunction.jl), written against the libsgp C ABI exactly as the Julia shim in INTEGRATION.md is.

The reference's rules run once per data point and ReactiveMP folds their messages with ``prod``; the seam the
reference itself provides is ``BufferUniSGP`` + ``meta.counter / meta.N`` (UniSGPnode.jl:56-73): the N-th ``prod`` is
where the expensive step fires.  The mirror keeps every rule signature (q_out, q_in, q_v, q_w, q_theta, meta):

* a per-point rule only *enqueues* (x_n, E[y_n], Var[y_n]) into the meta's staging buffers and returns a neutral
  message -- an untouched Gaussian for ``:v`` (``prod`` returns ``left``), ``GammaShapeRate(1, 0)`` for ``:w``, 0 for
  the average energy;
* the N-th call flushes: one ``sgp_set_data`` + ``sgp_sweep_psi`` (+ ``sgp_posterior_v`` / ``sgp_w_terms``) and returns
  the message whose product with the neutral ones equals the product of the reference's N per-point messages.

The marginals the host sees after each full pass are therefore identical to the reference's (tests/test_nodes_gpu.py
checks them against the per-point oracle)."""
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional

import numpy as np
from scipy.special import digamma

from .sgp import SGPContext

LOG2PI = float(np.log(2.0 * np.pi))


# ---- the few distribution types the rules exchange (ReactiveMP names) ---------------------------------------------
@dataclass
class PointMass:
    value: Any


@dataclass
class NormalMeanVariance:
    m: float
    v: float


@dataclass
class NormalMeanPrecision:
    m: float
    w: float


@dataclass
class GammaShapeRate:
    a: float
    b: float


@dataclass
class MvNormalMeanCovariance:
    m: np.ndarray
    S: np.ndarray


@dataclass
class MvNormalMeanPrecision:
    m: np.ndarray
    W: np.ndarray


@dataclass
class MvNormalWeightedMeanPrecision:
    xi: np.ndarray
    Lam: np.ndarray
    _mean: Optional[np.ndarray] = None
    _cov: Optional[np.ndarray] = None


@dataclass
class Wishart:
    nu: float
    S: np.ndarray        # scale: E[W] = nu S


@dataclass
class WishartFast:
    nu: float
    invS: np.ndarray     # ReactiveMP's inverse-scale parametrisation (MultiSGPnode.jl:404)


def mean(q):
    if isinstance(q, PointMass):
        return q.value
    if isinstance(q, (NormalMeanVariance, NormalMeanPrecision, MvNormalMeanCovariance, MvNormalMeanPrecision)):
        return q.m
    if isinstance(q, GammaShapeRate):
        return q.a / q.b
    if isinstance(q, Wishart):
        return q.nu * q.S
    if isinstance(q, MvNormalWeightedMeanPrecision):
        return mean_cov(q)[0]
    raise TypeError(type(q))


def mean_var(q):
    if isinstance(q, PointMass):
        return q.value, 0.0
    if isinstance(q, NormalMeanVariance):
        return q.m, q.v
    if isinstance(q, NormalMeanPrecision):
        return q.m, 1.0 / q.w
    raise TypeError(type(q))


def mean_cov(q):
    if isinstance(q, MvNormalMeanCovariance):
        return q.m, q.S
    if isinstance(q, MvNormalMeanPrecision):
        return q.m, np.linalg.inv(q.W)
    if isinstance(q, MvNormalWeightedMeanPrecision):
        if q._mean is None:       # only reached for messages this module did not produce
            q._cov = np.linalg.inv(q.Lam)
            q._mean = q._cov @ q.xi
        return q._mean, q._cov
    raise TypeError(type(q))


def mean_log(q):
    if isinstance(q, GammaShapeRate):
        return float(digamma(q.a) - np.log(q.b))
    if isinstance(q, PointMass):
        return float(np.log(q.value))
    raise TypeError(type(q))


# ---- meta ---------------------------------------------------------------------------------------------------------
@dataclass
class UniSGPMeta:
    """helper_functions/gp_helperfunction.jl:33-44 -- same fields, same order -- plus the library handle and the staging
    buffers the per-point rules fill.  ``kernel(theta)`` returns (variance, lengthscale[, kind])."""
    method: Any
    Xu: np.ndarray
    Psi0: np.ndarray
    Psi1_trans: np.ndarray
    Psi2: np.ndarray
    KuuL: Any
    kernel: Callable
    Uv: Any
    counter: int
    N: int
    ctx: Optional[SGPContext] = None
    kuu_jitter: float = 0.0
    _x: List = field(default_factory=list)
    _y: List = field(default_factory=list)
    _yv: List = field(default_factory=list)
    _wcount: int = 0
    _ecount: int = 0
    _theta_key: Any = None
    _swept: bool = False


def getmethod(meta): return meta.method
def getInducingInput(meta): return meta.Xu
def getKernel(meta): return meta.kernel
def getUv(meta): return meta.Uv


@dataclass
class BufferUniSGP:
    """GPnode/UniSGPnode.jl:56-59."""
    qv: Any
    meta: UniSGPMeta


def _ctx(meta):
    if meta.ctx is None:
        meta.ctx = SGPContext(0)     # raises without a GPU / library: no CPU fallback
    return meta.ctx


def _configure(meta, theta):
    """kernel(theta) + inducing inputs into the library; refactor K_uu when theta changed (host-side in the reference:
    experiments/regression_kin40k.ipynb:183-184)."""
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    key = theta.tobytes()
    if meta._theta_key == key:
        return
    ctx = _ctx(meta)
    k = meta.kernel(theta)
    var, ell = k[0], k[1]
    kind = k[2] if len(k) > 2 else 0
    Z = np.asarray(meta.Xu, dtype=np.float64)
    Z = Z[:, None] if Z.ndim == 1 else Z
    ctx.set_kernel(var, ell, D=Z.shape[1], kind=kind)
    ctx.set_inducing(Z)
    meta._theta_key = key
    meta._swept = False


def _enqueue(meta, q_out, q_in):
    mu_y, v_y = mean_var(q_out)
    meta._x.append(np.atleast_1d(np.asarray(mean(q_in), dtype=np.float64)))
    meta._y.append(float(mu_y)); meta._yv.append(float(v_y))


def _upload(meta):
    ctx = _ctx(meta)
    X = np.stack(meta._x[:meta.N]); y = np.array(meta._y[:meta.N]); yv = np.array(meta._yv[:meta.N])
    ctx.set_data(X, y, yv if np.any(yv != 0.0) else None)
    meta._x.clear(); meta._y.clear(); meta._yv.clear()


# ---- :v rule + prod -----------------------------------------------------------------------------------------------
def rule_v(q_out, q_in, q_w, q_theta, meta: UniSGPMeta):
    """@rule UniSGP(:v, Marginalisation) (q_out::PointMass|Gaussian, q_in::PointMass, q_w, q_theta, meta)
    -- GPnode/UniSGPnode.jl:144-158, 161-173.  Enqueues; the message is materialised by the N-th ``prod``."""
    _configure(meta, mean(q_theta))
    _enqueue(meta, q_out, q_in)
    return BufferUniSGP((float(mean(q_w)),), meta)


def prod(left, right: BufferUniSGP):
    """ReactiveMP.prod(::GenericProd, left::NormalDistributionsFamily, right::BufferUniSGP) -- UniSGPnode.jl:62-73.
    Same argument order, same counter / N semantics; on ``counter == N`` the sweep runs on the GPU, the posterior is
    factorised, ``meta.Uv`` refreshed and the counter reset."""
    meta = right.meta
    meta.counter += 1
    if meta.counter != meta.N:
        return left
    ctx = _ctx(meta)
    w = right.qv[0]
    _upload(meta)
    psi0, psi1, psi2, sy2 = ctx.sweep_psi()
    meta.Psi0[...] = psi0; meta.Psi1_trans[:, 0] = psi1; meta.Psi2[...] = psi2
    meta._swept = True
    if isinstance(left, MvNormalWeightedMeanPrecision):
        xi0, Lam0 = left.xi, left.Lam
    else:
        m0, S0 = mean_cov(left)
        Lam0 = np.linalg.inv(S0); xi0 = Lam0 @ m0
    mu_v, Sigma_v, Uv = ctx.posterior_v(xi0, Lam0, w)
    meta.Uv = Uv
    meta.counter = 0
    return MvNormalWeightedMeanPrecision(xi0 + w * psi1, Lam0 + w * psi2, mu_v, Sigma_v)


# ---- :w rule ------------------------------------------------------------------------------------------------------
def _ensure_kuu(meta):
    ctx = _ctx(meta)
    if meta.KuuL is None or getattr(meta, "_kuu_key", None) != meta._theta_key:
        meta.KuuL = ctx.kuu_factor(meta.kuu_jitter)
        meta._kuu_key = meta._theta_key


def _w_terms(meta, q_v):
    ctx = _ctx(meta)
    _ensure_kuu(meta)
    if len(meta._x) >= meta.N and meta.N > 0:    # this pass brought its own data
        _upload(meta)
        meta._swept = False
    if not meta._swept:
        ctx.sweep_psi(fetch=False)
        meta._swept = True
    return ctx.w_terms(mean(q_v), meta.Uv)


def rule_w(q_out, q_in, q_v, q_theta, meta: UniSGPMeta):
    """@rule UniSGP(:w, Marginalisation) -- UniSGPnode.jl:196-216, 219-238.  Neutral GammaShapeRate(1, 0) for the first
    N-1 nodes, GammaShapeRate(1 + N/2, sum_n rate_n) on the N-th: the product over the N nodes equals the reference's."""
    _configure(meta, mean(q_theta))
    _enqueue(meta, q_out, q_in)
    meta._wcount += 1
    if meta._wcount != meta.N:
        return GammaShapeRate(1.0, 0.0)
    meta._wcount = 0
    s1, s2 = _w_terms(meta, q_v)
    return GammaShapeRate(1.0 + 0.5 * meta.N, 0.5 * (s1 + s2))


def average_energy(q_out, q_in, q_v, q_w, q_theta, meta: UniSGPMeta):
    """@average_energy UniSGP -- UniSGPnode.jl:337-359, 363-387: 0 for the first N-1 nodes, sum_n U_n on the N-th."""
    _configure(meta, mean(q_theta))
    _enqueue(meta, q_out, q_in)
    meta._ecount += 1
    if meta._ecount != meta.N:
        return 0.0
    meta._ecount = 0
    s1, s2 = _w_terms(meta, q_v)
    w_bar = mean(q_w)
    return 0.5 * w_bar * (s1 + s2) + 0.5 * meta.N * (LOG2PI - mean_log(q_w))


# ---- :out rule ----------------------------------------------------------------------------------------------------
def rule_out(q_in, q_v, q_w, q_theta, meta: UniSGPMeta):
    """@rule UniSGP(:out, Marginalisation) (q_in::PointMass, ...) -- UniSGPnode.jl:96-104; ``q_in`` may hold a whole
    test set (rows), which is how experiments/regression_kin40k.ipynb:289-304 uses it in a loop."""
    _configure(meta, mean(q_theta))
    x = np.asarray(mean(q_in), dtype=np.float64)
    D = _ctx(meta).D
    X = x.reshape(-1, D)
    m = _ctx(meta).predict_mean(X, mean(q_v))
    w = mean(q_w)
    if X.shape[0] == 1 and x.ndim <= 1:
        return NormalMeanPrecision(float(m[0]), w)
    return [NormalMeanPrecision(float(v), w) for v in m]
