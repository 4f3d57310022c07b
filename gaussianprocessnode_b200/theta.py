"""Host-side mirror of the theta-step helpers (helper_functions/derivative_helper.jl) over sgp_theta_objective.

Same names and keyword arguments as the reference: ``neg_log_backwardmess_fast(theta; y_data, x_data, v, Uv, w, kernel,
Xu)`` (derivative_helper.jl:23-39) and ``grad_llh_new!(grad, theta; ..., chunk_size)`` (:59-63; ``chunk_size`` is a ForwardDiff
tuning knob and is ignored).  ``kernel(theta)`` returns ``(variance, lengthscale[, kind])`` as everywhere in this mirror; the
gradient with respect to the raw theta needs the Jacobian of that map: ``kernel_jac(theta)`` -> (d variance / d theta [P],
d lengthscale / d theta [D x P]); when omitted, the notebooks' parametrisation ``softplus.(theta)`` with
theta = [variance_raw, lengthscale_raw...] is assumed (experiments/regression_kin40k.ipynb:108)."""
import numpy as np

from .sgp import SGPContext


def _softplus_jac(theta, D):
    s = 1.0 / (1.0 + np.exp(-np.asarray(theta, dtype=np.float64)))          # softplus'
    P = theta.size
    dvar = np.zeros(P); dvar[0] = s[0]
    dell = np.zeros((D, P))
    if P == 2 and D > 1:                                                    # one shared lengthscale
        dell[:, 1] = s[1]
    else:
        for d in range(D):
            dell[d, 1 + d] = s[1 + d]
    return dvar, dell


def _load(ctx, theta, y_data, x_data, kernel, Xu, meta=None):
    if meta is not None:          # the node meta's context is being re-pointed: its cached kernel / data / sweep state no longer describes the library
        meta._theta_key = None; meta._resident_sig = None; meta._swept = False
        if hasattr(meta, "_kuu_key"):
            meta._kuu_key = None
    k = kernel(np.asarray(theta, dtype=np.float64))
    var, ell = k[0], k[1]
    kind = k[2] if len(k) > 2 else 0
    Z = np.asarray(Xu, dtype=np.float64); Z = Z[:, None] if Z.ndim == 1 else Z
    X = np.asarray(x_data, dtype=np.float64); X = X[:, None] if X.ndim == 1 else X
    ctx.set_kernel(var, ell, D=Z.shape[1], kind=kind); ctx.set_inducing(Z); ctx.set_data(X, np.asarray(y_data, dtype=np.float64))
    return Z.shape[1]


def neg_log_backwardmess_fast(theta, *, y_data, x_data, v, Uv, w, kernel, Xu, ctx=None, jitter=0.0, meta=None):
    """derivative_helper.jl:23-39 -- the value of the collapsed objective.  `meta`: a node meta whose context is used (and whose cached
    state is invalidated, because this call re-points the context's kernel and data)."""
    if meta is not None and ctx is None:
        from .nodes import _ctx
        ctx = _ctx(meta)
    own = ctx is None
    ctx = SGPContext(0) if own else ctx
    try:
        _load(ctx, theta, y_data, x_data, kernel, Xu, meta)
        return ctx.theta_objective(v, Uv, w, jitter, grad=False)
    finally:
        if own:
            ctx.close()


def grad_llh_new(grad, theta, *, y_data, x_data, v, Uv, w, kernel, Xu, chunk_size=None, kernel_jac=None, ctx=None, jitter=0.0, meta=None):
    """grad_llh_new!(grad, theta; ...) -- derivative_helper.jl:59-63: fills ``grad`` with d(objective)/d(theta) and returns it."""
    if meta is not None and ctx is None:
        from .nodes import _ctx
        ctx = _ctx(meta)
    own = ctx is None
    ctx = SGPContext(0) if own else ctx
    try:
        theta = np.asarray(theta, dtype=np.float64)
        D = _load(ctx, theta, y_data, x_data, kernel, Xu, meta)
        _, dvar, dell = ctx.theta_objective(v, Uv, w, jitter, grad=True)
        jv, jl = kernel_jac(theta) if kernel_jac is not None else _softplus_jac(theta, D)
        grad[...] = dvar * np.asarray(jv) + np.asarray(dell) @ np.asarray(jl)
        return grad
    finally:
        if own:
            ctx.close()
