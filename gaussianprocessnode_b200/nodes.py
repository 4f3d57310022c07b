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
    _q: Any = field(default_factory=lambda: {"v": [], "w": [], "e": []})    # one staging queue per interface: rule families may interleave
    _theta_key: Any = None
    _swept: bool = False
    _resident_sig: Any = None                                                 # what the library currently holds (points, targets)


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
    meta._resident_sig = None


def _enqueue(meta, kind, q_out, q_in):
    """One (x_n, E[y_n], Var[y_n]) into the staging queue of interface `kind` ("v", "w", "e"); True when the queue holds meta.N entries."""
    mu_y, v_y = mean_var(q_out)
    q = meta._q[kind]
    q.append((np.atleast_1d(np.asarray(mean(q_in), dtype=np.float64)), float(mu_y), float(v_y)))
    return len(q) == meta.N


def _upload(meta, kind):
    """Flushes the queue of interface `kind` into the library (sgp_set_data); a pass that queued exactly what is already resident
    (the :w rule / energy after the :v rule of the same iteration) costs nothing."""
    ctx = _ctx(meta)
    q = meta._q[kind]
    assert len(q) == meta.N, "UniSGP %s interface: %d queued points, meta.N = %d" % (kind, len(q), meta.N)
    X = np.stack([e[0] for e in q]); y = np.array([e[1] for e in q]); yv = np.array([e[2] for e in q])
    q.clear()
    sig = (meta._theta_key, X.tobytes(), y.tobytes(), yv.tobytes())
    if sig == meta._resident_sig:
        return
    ctx.set_data(X, y, yv if np.any(yv != 0.0) else None)
    meta._resident_sig = sig
    meta._swept = False


# ---- :v rule + prod -----------------------------------------------------------------------------------------------
def rule_v(q_out, q_in, q_w, q_theta, meta: UniSGPMeta):
    """@rule UniSGP(:v, Marginalisation) (q_out::PointMass|Gaussian, q_in::PointMass, q_w, q_theta, meta)
    -- GPnode/UniSGPnode.jl:144-158, 161-173.  Enqueues; the message is materialised by the N-th ``prod``."""
    _configure(meta, mean(q_theta))
    _enqueue(meta, "v", q_out, q_in)
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
    _upload(meta, "v")
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


def _w_terms(meta, kind, q_v):
    ctx = _ctx(meta)
    _ensure_kuu(meta)
    _upload(meta, kind)                          # this pass brought its own data (a no-op when it is what the :v pass left resident)
    if not meta._swept:
        ctx.sweep_psi(fetch=False)
        meta._swept = True
    return ctx.w_terms(mean(q_v), meta.Uv)


def rule_w(q_out, q_in, q_v, q_theta, meta: UniSGPMeta):
    """@rule UniSGP(:w, Marginalisation) -- UniSGPnode.jl:196-216, 219-238.  Neutral GammaShapeRate(1, 0) for the first
    N-1 nodes, GammaShapeRate(1 + N/2, sum_n rate_n) on the N-th: the product over the N nodes equals the reference's."""
    _configure(meta, mean(q_theta))
    if not _enqueue(meta, "w", q_out, q_in):
        return GammaShapeRate(1.0, 0.0)
    s1, s2 = _w_terms(meta, "w", q_v)
    return GammaShapeRate(1.0 + 0.5 * meta.N, 0.5 * (s1 + s2))


def average_energy(q_out, q_in, q_v, q_w, q_theta, meta: UniSGPMeta):
    """@average_energy UniSGP -- UniSGPnode.jl:337-359, 363-387: 0 for the first N-1 nodes, sum_n U_n on the N-th."""
    _configure(meta, mean(q_theta))
    if not _enqueue(meta, "e", q_out, q_in):
        return 0.0
    s1, s2 = _w_terms(meta, "e", q_v)
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


# ======================================================================================================================
# Uncertain inputs: UniSGP with q(x) Gaussian (GPnode/UniSGPnode.jl:125-140) and the MultiSGP node
# (GPnode/MultiSGPnode.jl, helper_functions/gp_helperfunction.jl:55-73)
# ======================================================================================================================
def _method_of(meta):
    """meta.method: a libsgp method id or (id, p) for Gauss-Hermite of order p (the reference passes a ReactiveMP
    approximation object: srcubature(), ghcubature(p), GenUnscented())."""
    return meta.method if isinstance(meta.method, tuple) else (meta.method, 21)


def rule_v_uncertain(q_out, q_in, q_w, q_theta, meta: UniSGPMeta):
    """@rule UniSGP(:v, Marginalisation) (q_out, q_in::UnivariateGaussianDistributionsFamily, ...) -- UniSGPnode.jl:125-140.
    Enqueues (m_n, v_n, E[y_n]); ``prod_uncertain`` materialises the sum on the N-th fold."""
    _configure(meta, mean(q_theta))
    m, v = mean_var(q_in)
    mu_y, _ = mean_var(q_out)
    meta._q["v"].append((np.atleast_1d(np.asarray(m, dtype=np.float64)), np.atleast_2d(np.asarray(v, dtype=np.float64)), float(mu_y)))
    return BufferUniSGP((float(mean(q_w)), "uncertain"), meta)


def prod_uncertain(left, right: BufferUniSGP):
    """The N-th fold of the uncertain-input messages (the same `prod` UniSGPnode.jl:62-73 serves every :v variant): one
    sgp_sweep_psi_uncertain over the N queued inputs, then the posterior exactly as `prod` does -- mean / covariance of the folded Gaussian
    and ``meta.Uv = chol(Sigma_v + mu_v mu_v').U``.  Each per-node message carries Psi2_n + 1e-8 I (:135), so the sum carries N * 1e-8 I."""
    meta = right.meta
    meta.counter += 1
    if meta.counter != meta.N:
        return left
    ctx = _ctx(meta)
    w = right.qv[0]
    mid, p = _method_of(meta)
    q = meta._q["v"]
    assert len(q) == meta.N, "UniSGP :v interface (uncertain inputs): %d queued nodes, meta.N = %d" % (len(q), meta.N)
    means = np.stack([e[0] for e in q]); covs = np.stack([e[1] for e in q]); y = np.array([e[2] for e in q])
    q.clear()
    meta._resident_sig = None; meta._swept = False
    psi0, psi1, psi2, _ = ctx.sweep_psi_uncertain(mid, means, covs, R=y[:, None], D_out=1, p=p)
    psi1 = np.asarray(psi1).reshape(-1)
    jit = meta.N * 1e-8 * np.eye(psi2.shape[0])
    meta.Psi0[...] = psi0; meta.Psi1_trans[:, 0] = psi1; meta.Psi2[...] = psi2 + jit
    if isinstance(left, MvNormalWeightedMeanPrecision):
        xi0, Lam0 = left.xi, left.Lam
    else:
        m0, S0 = mean_cov(left)
        Lam0 = np.linalg.inv(S0); xi0 = Lam0 @ m0
    # the resident statistics are the un-jittered sums: the jitter rides on the prior's precision
    mu_v, Sigma_v, Uv = ctx.posterior_v(xi0, Lam0 + w * jit, w)
    meta.Uv = Uv
    meta.counter = 0
    return MvNormalWeightedMeanPrecision(xi0 + w * psi1, Lam0 + w * (psi2 + jit), mu_v, Sigma_v)


def _uncertain_node_terms(q_outs, q_ins, q_v, q_theta, meta: UniSGPMeta):
    """Per-node I1_n, I2_n of the uncertain-input UniSGP rules (UniSGPnode.jl:177-192, 290-313) for a whole set of nodes: the sigma-point
    cloud of ONE sgp_sweep_psi_uncertain, then sgp_uncertain_node_terms (kernel columns of all sigma points -> K_uu^-1 K and U_v K tile
    GEMMs -> per-node weighted sums).  The reference adds 1e-8 I to every Psi2_n and clamps both terms to [1e-12, 1e12] per node."""
    _configure(meta, mean(q_theta))
    ctx = _ctx(meta)
    _ensure_kuu(meta)
    mid, p = _method_of(meta)
    mv = [mean_var(q) for q in q_ins]
    means = np.stack([np.atleast_1d(np.asarray(m, dtype=np.float64)) for m, _ in mv])
    covs = np.stack([np.atleast_2d(np.asarray(v, dtype=np.float64)) for _, v in mv])
    N = means.shape[0]
    ctx.sweep_psi_uncertain(mid, means, covs, R=None, D_out=1, p=p)          # builds (and leaves resident) the sigma-point cloud
    meta._swept = False                                                      # the resident statistics are now the cloud's
    psi0, qk, lin, qr, tr_kinv, frob_uv = ctx.uncertain_node_terms(mean(q_v), meta.Uv, N)
    I1 = np.clip(psi0 - qk - 1e-8 * tr_kinv, 1e-12, 1e12)
    if q_outs is None:
        return I1, None, lin
    yv = [mean_var(q) for q in q_outs]
    y = np.array([float(a) for a, _ in yv]); vy = np.array([float(b) for _, b in yv])
    I2 = np.clip(y * y + vy - 2.0 * y * lin + qr + 1e-8 * frob_uv, 1e-12, 1e12)
    return I1, I2, lin


def rule_w_uncertain(q_outs, q_ins, q_v, q_theta, meta: UniSGPMeta):
    """@rule UniSGP(:w, Marginalisation) (q_out, q_in::UnivariateGaussianDistributionsFamily, ...) -- UniSGPnode.jl:177-192 -- for a whole set
    of nodes: the N messages GammaShapeRate(1.5, (I1_n + I2_n) / 2) (the clamps are per node, so the messages are returned individually)."""
    I1, I2, _ = _uncertain_node_terms(q_outs, q_ins, q_v, q_theta, meta)
    return [GammaShapeRate(1.5, 0.5 * float(a + b)) for a, b in zip(I1, I2)]


def rule_out_uncertain(q_ins, q_v, q_w, q_theta, meta: UniSGPMeta):
    """@rule UniSGP(:out, Marginalisation) (q_in::UnivariateGaussianDistributionsFamily, ...) -- UniSGPnode.jl:85-93: NormalMeanPrecision(Psi1_n' mu_v, w_bar)."""
    _, _, lin = _uncertain_node_terms(None, q_ins, q_v, q_theta, meta)
    w = mean(q_w)
    return [NormalMeanPrecision(float(v), w) for v in lin]


def average_energy_uncertain(q_outs, q_ins, q_v, q_w, q_theta, meta: UniSGPMeta):
    """@average_energy UniSGP (q_out, q_in::UnivariateGaussianDistributionsFamily, ..., q_w::GammaShapeRate) -- UniSGPnode.jl:290-313: the N
    per-node energies U_n = (I1_n w_bar - E[ln w] + ln 2 pi + I2_n w_bar) / 2."""
    I1, I2, _ = _uncertain_node_terms(q_outs, q_ins, q_v, q_theta, meta)
    w_bar = mean(q_w)
    return 0.5 * (I1 * w_bar - mean_log(q_w) + LOG2PI + I2 * w_bar)


def _wpoint_energies(q_outs, q_ins, q_v, q_w, q_theta, meta: UniSGPMeta, pointmass_in):
    """The two `q_w::PointMass` energies that add 1e-8 to EVERY element of an un-jittered K_uu, of Psi1_n and of Psi2_n and call plain `inv`
    (UniSGPnode.jl:390-409 with q_in Gaussian, :438-458 with q_in PointMass), per node, on the library:
      inv(K + eps 1 1') = K^-1 - c u u',  u = K^-1 1,  c = eps / (1 + eps 1'u)                      (Sherman-Morrison on the Cholesky-based K^-1)
      tr(inv(.) (Psi2_n + eps 1 1')) = tr(K^-1 Psi2_n) - c u' Psi2_n u + eps (1'u - c (1'u)^2)
      tr(R_v (Psi2_n + eps 1 1'))    = tr(R_v Psi2_n) + eps 1' R_v 1,      (Psi1_n + eps 1)' mu_v = Psi1_n' mu_v + eps 1' mu_v
    with the per-node traces from sgp_uncertain_node_terms (u' Psi2_n u is the same call with the "factor" [u'; 0]).  K_uu without any jitter
    is near-singular, so the value of the reference's LU-based `inv` is defined only up to cond(K_uu) * eps (tests/test_nodes_gpu.py states
    the spread)."""
    _configure(meta, mean(q_theta))
    ctx = _ctx(meta)
    eps = 1e-8
    ctx.kuu_factor(0.0, fetch=False)
    meta.KuuL = None; meta._kuu_key = None                    # the library now holds the UN-jittered factor, not meta.kuu_jitter's
    M = np.asarray(meta.Xu).shape[0]
    Kinv = ctx.kuu_solve(np.eye(M))
    u = Kinv @ np.ones(M); s = float(u.sum()); c = eps / (1.0 + eps * s)
    mu_v, Sigma_v = mean_cov(q_v)
    Rv = np.asarray(Sigma_v, dtype=np.float64) + np.outer(mu_v, mu_v)
    Uv = np.linalg.cholesky(Rv).T
    if pointmass_in:
        from .sgp import POINT
        means = np.stack([np.atleast_1d(np.asarray(mean(q), dtype=np.float64)) for q in q_ins]); covs = None; mid, p = POINT, 0
    else:
        mid, p = _method_of(meta)
        mv = [mean_var(q) for q in q_ins]
        means = np.stack([np.atleast_1d(np.asarray(m, dtype=np.float64)) for m, _ in mv])
        covs = np.stack([np.atleast_2d(np.asarray(v, dtype=np.float64)) for _, v in mv])
    N = means.shape[0]
    meta._resident_sig = None; meta._swept = False
    ctx.sweep_psi_uncertain(mid, means, covs, R=None, D_out=1, p=p)
    psi0, qk, lin, qr, _, _ = ctx.uncertain_node_terms(mu_v, Uv, N)
    U2 = np.zeros((M, M)); U2[0, :] = u
    _, _, _, quu, _, _ = ctx.uncertain_node_terms(mu_v, U2, N)
    yv = [mean_var(q) for q in q_outs]
    y = np.array([float(a) for a, _ in yv]); vy = np.array([float(b) for _, b in yv])
    I1 = np.clip(psi0 - (qk - c * quu) - eps * (s - c * s * s), 1e-12, 1e12)
    I2 = np.clip(y * y + vy - 2.0 * y * (lin + eps * float(np.sum(mu_v))) + qr + eps * float(Rv.sum()), 1e-12, 1e12)
    w_bar = float(mean(q_w))
    return 0.5 * (I1 * w_bar - np.log(w_bar) + LOG2PI + I2 * w_bar)


def average_energy_uncertain_wpoint(q_outs, q_ins, q_v, q_w: PointMass, q_theta, meta: UniSGPMeta):
    """@average_energy UniSGP (q_out Gaussian, q_in::UnivariateGaussianDistributionsFamily, q_v, q_w::PointMass, ...) -- UniSGPnode.jl:390-409: the N
    per-node energies."""
    return _wpoint_energies(q_outs, q_ins, q_v, q_w, q_theta, meta, False)


def average_energy_gaussout_wpoint(q_outs, q_ins, q_v, q_w: PointMass, q_theta, meta: UniSGPMeta):
    """@average_energy UniSGP (q_out Gaussian, q_in::PointMass, q_v, q_w::PointMass, ...) -- UniSGPnode.jl:438-458: the N per-node energies."""
    return _wpoint_energies(q_outs, q_ins, q_v, q_w, q_theta, meta, True)


def predict_new(x_test, q_v, q_w, q_theta, meta: UniSGPMeta):
    """`predict_new` of the classification drivers (experiments/classification_banana.ipynb:289-293): the :out message of every test point
    pushed through `@rule Probit(:out)`.  Returns (prediction_f: list of NormalMeanPrecision, p: Bernoulli means, p (1 - p): their variances)."""
    _configure(meta, mean(q_theta))
    ctx = _ctx(meta)
    X = np.asarray(x_test, dtype=np.float64).reshape(-1, ctx.D)
    w = float(mean(q_w))
    m, _, prob = ctx.predict_probit(X, mean(q_v), w)
    return [NormalMeanPrecision(float(v), w) for v in m], prob, prob * (1.0 - prob)


@dataclass
class MultiSGPMeta:
    """helper_functions/gp_helperfunction.jl:55-64 -- same fields, same order (GPCache is scratch only and has no
    counterpart) -- plus N (nodes sharing the meta), the library handle and the staging buffers."""
    method: Any
    Xu: np.ndarray
    Psi0: np.ndarray
    Psi1_trans: np.ndarray
    Psi2: np.ndarray
    Kuu_inverse: Any
    kernel: Callable
    N: int = 0
    ctx: Optional[SGPContext] = None
    _m: List = field(default_factory=list)
    _P: List = field(default_factory=list)
    _muy: List = field(default_factory=list)
    _Sy: List = field(default_factory=list)
    _count: int = 0
    _theta_key: Any = None
    _stats: Any = None


def _multi_enqueue(meta, q_out, q_in):
    m, P = mean_cov(q_in)
    if isinstance(q_out, PointMass):
        mu_y, S_y = np.asarray(q_out.value, dtype=np.float64), None
    else:
        mu_y, S_y = mean_cov(q_out)
    meta._m.append(np.asarray(m, dtype=np.float64)); meta._P.append(np.asarray(P, dtype=np.float64))
    meta._muy.append(np.asarray(mu_y, dtype=np.float64)); meta._Sy.append(S_y)
    meta._count += 1
    return meta._count == meta.N


def _multi_flush(meta, want_psi1_n=False):
    """One sgp_sweep_psi_uncertain over the N queued nodes: Psi0, P_Y = sum_n Psi1_n mu_y_n' (M x D), Psi2 (and Psi1_n)."""
    ctx = _ctx(meta)
    mid, p = _method_of(meta)
    means = np.stack(meta._m[:meta.N]); covs = np.stack(meta._P[:meta.N]); Y = np.stack(meta._muy[:meta.N])
    Sy = sum((np.zeros((Y.shape[1], Y.shape[1])) if s is None else s) for s in meta._Sy[:meta.N])
    meta._m.clear(); meta._P.clear(); meta._muy.clear(); meta._Sy.clear(); meta._count = 0
    psi0, PY, psi2, p1n = ctx.sweep_psi_uncertain(mid, means, covs, R=Y, D_out=Y.shape[1], p=p, want_psi1_n=want_psi1_n)
    PY = np.asarray(PY).reshape(psi2.shape[0], Y.shape[1])
    meta.Psi0[...] = psi0; meta.Psi2[...] = psi2
    meta._stats = dict(psi0=psi0, PY=PY, psi2=psi2, Y=Y, Sy=Sy, psi1_n=p1n)
    return meta._stats


def multi_rule_v(q_out, q_in, q_w, q_theta, meta: MultiSGPMeta):
    """@rule MultiSGP(:v, Marginalisation) -- MultiSGPnode.jl:290-308 (q_out Gaussian), :310-328 (PointMass).
    Neutral (xi = 0, Lambda = 0) for the first N-1 nodes; on the N-th the sum of the N per-node messages:
    Lambda = kron(W_bar, sum_n Psi2_n),  xi = vec(sum_n Psi1_n (W_bar mu_y_n)')  (output-major blocks of M)."""
    _configure(meta, mean(q_theta))
    W = np.asarray(mean(q_w), dtype=np.float64)
    M = np.asarray(meta.Xu).shape[0]; D = W.shape[0]
    if not _multi_enqueue(meta, q_out, q_in):
        return MvNormalWeightedMeanPrecision(np.zeros(D * M), np.zeros((D * M, D * M)))
    s = _multi_flush(meta)
    return MvNormalWeightedMeanPrecision((s["PY"] @ W).ravel(order="F"), np.kron(W, s["psi2"]))


def _kuu_inverse(meta):
    """meta.Kuu_inverse = cholinv(Kuu + jitter I) (host side in the reference: Pendulum_Wishart_2d.ipynb:2542-2543)."""
    if meta.Kuu_inverse is None or getattr(meta, "_kuu_key", None) != meta._theta_key:
        ctx = _ctx(meta)
        ctx.kuu_factor(getattr(meta, "kuu_jitter", 1e-12), fetch=False)
        M = np.asarray(meta.Xu).shape[0]
        meta.Kuu_inverse = ctx.kuu_solve(np.eye(M))
        meta._kuu_key = meta._theta_key
    return meta.Kuu_inverse


def multi_rule_w(q_out, q_in, q_v, q_theta, meta: MultiSGPMeta):
    """@rule MultiSGP(:w, Marginalisation) -- MultiSGPnode.jl:367-405, 407-444.  The reference returns
    WishartFast(D+2, Psi_4,n) per node; their product has nu = N (D+2) - (N-1)(D+1) = N + D + 1 and inverse scale
    sum_n Psi_4,n.  Neutral element WishartFast(D+1, 0) for the first N-1 nodes."""
    _configure(meta, mean(q_theta))
    mu_v, Sigma_v = mean_cov(q_v)
    M = np.asarray(meta.Xu).shape[0]; D = mu_v.size // M
    if not _multi_enqueue(meta, q_out, q_in):
        return WishartFast(D + 1.0, np.zeros((D, D)))
    N = meta.N
    s = _multi_flush(meta)
    Kinv = _kuu_inverse(meta)
    Rv = Sigma_v + np.outer(mu_v, mu_v)
    V = mu_v.reshape(D, M).T                                            # column d = mu_v^(d)
    Psi4 = np.array([[np.sum(Rv[i * M:(i + 1) * M, j * M:(j + 1) * M] * s["psi2"].T) for j in range(D)] for i in range(D)])
    A = s["PY"].T @ V                                                   # sum_n mu_y_n E_n',  E_n = V' Psi1_n
    Ry = s["Y"].T @ s["Y"] + s["Sy"]
    I1 = s["psi0"] - np.sum(Kinv * s["psi2"])                           # sum_n (Psi0_n - tr(Kuu^-1 Psi2_n))
    return WishartFast(N + D + 1.0, Psi4 + Ry - (A + A.T) + I1 * np.eye(D))


def multi_average_energy(q_out, q_in, q_v, q_w, q_theta, meta: MultiSGPMeta):
    """@average_energy MultiSGP -- MultiSGPnode.jl:544-571, 574-602, 604-631: 0 for the first N-1 nodes, sum_n U_n on the
    N-th.  q_w: Wishart (E[W] = nu S, E[ln|W|] = psi_D(nu/2) + D ln 2 + ln|S|) or PointMass."""
    _configure(meta, mean(q_theta))
    mu_v, Sigma_v = mean_cov(q_v)
    M = np.asarray(meta.Xu).shape[0]; D = mu_v.size // M
    if not _multi_enqueue(meta, q_out, q_in):
        return 0.0
    N = meta.N
    s = _multi_flush(meta)
    Kinv = _kuu_inverse(meta)
    if isinstance(q_w, PointMass):
        W = np.asarray(q_w.value, dtype=np.float64); ElogW = float(np.linalg.slogdet(W)[1])
    else:
        W = q_w.nu * q_w.S
        ElogW = float(sum(digamma(0.5 * (q_w.nu - i)) for i in range(D)) + D * np.log(2.0) + np.linalg.slogdet(q_w.S)[1])
    Rv = Sigma_v + np.outer(mu_v, mu_v)
    V = mu_v.reshape(D, M).T
    sumRvblk_W = sum(Rv[i * M:(i + 1) * M, j * M:(j + 1) * M] * W[i, j] for i in range(D) for j in range(D))
    Ry = s["Y"].T @ s["Y"] + s["Sy"]
    return float(N * (0.5 * D * LOG2PI - 0.5 * ElogW) + 0.5 * np.trace(W @ Ry)
                 + 0.5 * np.trace(W) * (s["psi0"] - np.sum(Kinv * s["psi2"]))
                 - np.trace(W @ V.T @ s["PY"]) + 0.5 * np.sum(s["psi2"] * sumRvblk_W))


def multi_rule_out(q_ins, q_v, q_w, q_theta, meta: MultiSGPMeta):
    """@rule MultiSGP(:out, Marginalisation) -- MultiSGPnode.jl:90-104, 106-120 -- for a whole set of input marginals at
    once: MvNormalMeanPrecision([Psi1_n' mu_v^(d)]_d, W_bar) per node."""
    _configure(meta, mean(q_theta))
    ctx = _ctx(meta)
    mid, p = _method_of(meta)
    mc = [mean_cov(q) for q in q_ins]
    means = np.stack([np.asarray(m, dtype=np.float64) for m, _ in mc]); covs = np.stack([np.asarray(P, dtype=np.float64) for _, P in mc])
    mu_v = np.asarray(mean(q_v), dtype=np.float64)
    W = np.asarray(mean(q_w), dtype=np.float64)
    M = np.asarray(meta.Xu).shape[0]; D = mu_v.size // M
    _, _, _, p1n = ctx.sweep_psi_uncertain(mid, means, covs, p=p, want_psi1_n=True)
    E = p1n @ mu_v.reshape(D, M).T                                      # N x D
    return [MvNormalMeanPrecision(E[n], W) for n in range(E.shape[0])]


# ======================================================================================================================
# Driver-side helpers of helper_functions/gp_helperfunction.jl:133-158 (host code; no GPU work)
# ======================================================================================================================
def create_blockmatrix(A, d, M):
    """gp_helperfunction.jl:133-135: d x d grid of M x M views."""
    return [[A[i * M:(i + 1) * M, j * M:(j + 1) * M] for j in range(d)] for i in range(d)]


def split2batch(data, batch_size):
    """gp_helperfunction.jl:137-142: consecutive mini-batches, the last one ragged."""
    x, y = data
    n = len(x)
    return ([x[i:min(i + batch_size, n)] for i in range(0, n, batch_size)],
            [y[i:min(i + batch_size, n)] for i in range(0, len(y), batch_size)])


def SMSE(y_true, y_approx):
    """gp_helperfunction.jl:145-149: mean squared error over var(y_true), Julia's var = n-1 normalisation."""
    y_true = np.asarray(y_true, dtype=np.float64); y_approx = np.asarray(y_approx, dtype=np.float64)
    return float(np.sum((y_true - y_approx) ** 2) / y_true.size / np.var(y_true, ddof=1))


def num_error(ytrue, y):
    """gp_helperfunction.jl:152-154."""
    return float(np.sum(np.abs(np.asarray(y, dtype=np.float64) - np.asarray(ytrue, dtype=np.float64))))


def error_rate(ytrue, y):
    """gp_helperfunction.jl:156-158."""
    return num_error(ytrue, y) / len(ytrue)


# ---- @rule MultiSGP(:in) ---------------------------------------------------------------------------------
def _in_operands(q_v, q_w, meta: MultiSGPMeta, D: int):
    """What every node's :in message shares (MultiSGPnode.jl:175-181): Mv = [mu_v^(1) ... mu_v^(D)] (M x D), S = sumRvblk_W, tr(W)."""
    W = np.asarray(mean(q_w), dtype=np.float64)
    mu_v, Sigma_v = mean_cov(q_v)
    M = np.asarray(meta.Xu).shape[0]
    Rv = np.asarray(Sigma_v, dtype=np.float64) + np.outer(mu_v, mu_v)
    S = sum(Rv[i * M:(i + 1) * M, j * M:(j + 1) * M] * W[i, j] for i in range(D) for j in range(D))
    Mv = np.asarray(mu_v, dtype=np.float64).reshape(D, M).T
    return W, Mv, S


def multi_rule_in(q_outs, q_v, q_w, q_theta, meta: MultiSGPMeta):
    """@rule MultiSGP(:in, Marginalisation) -- MultiSGPnode.jl:162-185 (q_out Gaussian), :187-211 (PointMass) -- for a whole chain of
    nodes: returns `logpdf(Xp)`, the batched counterpart of the N closures `log_backwardmess`: Xp (N, P, d) -> (N, P) values
    (with grad=True / hess=True also the analytic derivatives), evaluated by sgp_in_logmessage in one call."""
    _configure(meta, mean(q_theta))
    ctx = _ctx(meta)
    Y = np.stack([np.asarray(q.value if isinstance(q, PointMass) else mean(q), dtype=np.float64) for q in q_outs])
    W, Mv, S = _in_operands(q_v, q_w, meta, Y.shape[1])
    R = Y @ W                                                     # row n = (W mu_y,n)'  (W symmetric)
    ctx.kuu_factor(getattr(meta, "kuu_jitter", 1e-12), fetch=False)        # Pendulum_Wishart_2d.ipynb:2542-2543: cholinv(Kuu + 1e-12 I)
    trW = float(np.trace(W))

    def logpdf(Xp, grad=False, hess=False):
        return ctx.in_logmessage(np.asarray(Xp, dtype=np.float64), Mv, S, trW, R=R, grad=grad, hess=hess)
    return logpdf


def multi_prod_gaussian_logpdf(lefts, logpdf):
    """`ReactiveMP.prod(::GenericProd, left::MvGaussian, right::ContinuousMultivariateLogPdf)` (MultiSGPnode.jl:38-45) for N (left_n,
    right_n) pairs at once: the 2d+1 spherical-radial points of every left_n go through ONE sgp_in_logmessage call; the moment
    matching (2d+1 terms per node) is host arithmetic.  Returns (means (N, d), covs (N, d, d)); a node whose mean is NaN keeps left_n."""
    ms = np.stack([np.asarray(mean_cov(q)[0], dtype=np.float64) for q in lefts])
    Ps = np.stack([np.asarray(mean_cov(q)[1], dtype=np.float64) for q in lefts])
    N, d = ms.shape
    L = np.linalg.cholesky(Ps)
    c = np.sqrt(d + 1.0)
    # srcubature: centre with weight 1/(d+1), m +- sqrt(d+1) L e_j with weight 1/(2(d+1))
    pts = np.concatenate([ms[:, None, :], ms[:, None, :] + c * np.swapaxes(L, 1, 2), ms[:, None, :] - c * np.swapaxes(L, 1, 2)], axis=1)
    wts = np.concatenate([[1.0 / (d + 1.0)], np.full(2 * d, 0.5 / (d + 1.0))])
    g = np.exp(logpdf(pts)) * wts[None, :]
    Zn = g.sum(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        mu = np.einsum("np,npd->nd", g, pts) / Zn[:, None]
        dl = pts - mu[:, None, :]
        cov = np.einsum("np,npi,npj->nij", g, dl, dl) / Zn[:, None, None]
    bad = np.isnan(mu[:, 0])
    mu[bad] = ms[bad]; cov[bad] = Ps[bad]
    return mu, cov


def multi_rule_in_laplace(q_outs, q_ins, q_v, q_w, q_theta, meta: MultiSGPMeta, iterations=50, tol=1e-12):
    """@rule MultiSGP(:in) -- MultiSGPnode.jl:213-236: mode m_z of every node's backward message from mean(q_in) and W_z = Hessian of
    the negative log message there; returns (xi (N, d) = W_z m_z, W_z (N, d, d), m_z).  The reference runs LBFGS (20 iterations,
    ForwardDiff gradient) per node and Zygote.hessian; here all N nodes advance together by damped Newton steps on the analytic
    gradient / Hessian of sgp_in_logmessage (a step is halved until the message does not decrease; a non-positive-definite Hessian
    falls back to a gradient step)."""
    logpdf = multi_rule_in(q_outs, q_v, q_w, q_theta, meta)
    x = np.stack([np.asarray(mean_cov(q)[0], dtype=np.float64) for q in q_ins])
    N, d = x.shape
    f, g, H = logpdf(x[:, None, :], hess=True)
    f = f[:, 0]; g = g[:, 0]; H = H[:, 0]
    for _ in range(iterations):
        step = np.empty_like(x)
        for n in range(N):
            try:
                np.linalg.cholesky(-H[n])
                step[n] = np.linalg.solve(-H[n], g[n])
            except np.linalg.LinAlgError:
                step[n] = g[n] / max(np.linalg.norm(H[n], 2), 1e-12)
        t = np.ones(N)
        active = np.linalg.norm(g, axis=1) > tol
        if not active.any():
            break
        for _ls in range(30):
            xn = x + (t * active)[:, None] * step
            fn, gn, Hn = logpdf(xn[:, None, :], hess=True)
            fn = fn[:, 0]
            worse = active & ~(fn >= f - 1e-14 * np.abs(f))
            if not worse.any():
                break
            t[worse] *= 0.5
        x, f, g, H = xn, fn, gn[:, 0], Hn[:, 0]
    Wz = -H
    return np.einsum("nij,nj->ni", Wz, x), Wz, x
