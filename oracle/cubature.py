"""Sigma-point rules used for uncertain inputs (oracle; test infrastructure only).

* ``srcubature`` / ``ghcubature``: ReactiveMP.jl (NOT under /root/reference, unpinned; **parity unpinned**,
  restated from ReactiveMP's published definitions; call sites GPnode/MultiSGPnode.jl:15-35,
  GPnode/UniSGPnode.jl:11-33, GPtest.jl:14-15).  GPtest.jl:380 asserts the spherical-radial weights sum to one
  exactly for a constant integrand (tests/test_oracle_rules.py checks that).
  Searched for a stronger pin (round 2): every code cell of experiments/*.ipynb that touches srcubature / ghcubature /
  approximate_kernel_expectation / GenUnscented (Pendulum_Wishart_2d cells 13, 36, 39; GPLVM cell 14) either prints nothing or
  depends on Julia's MersenneTwister stream (`Random.seed!(124)` data, `rand` initialisations), and GPtest.jl draws its inputs
  with unseeded `rand`: the reference holds NO RNG-free cubature output, so node placement and order stay pinned only to the
  Monte-Carlo tolerances of GPtest.jl:141-143, 380-382 and to the published definitions.  Nothing further can be pinned here.
* ``gen_unscented_*``: helper_functions/ut_approx.jl:116-151, quirks reproduced as written.
All rules return (points [S x d], weights [S]) in the reference's enumeration order.
"""
import numpy as np


def cholsqrt(P):
    """FastCholesky.cholsqrt = lower Cholesky factor (unvendored; parity unpinned beyond 'is a Cholesky')."""
    P = np.atleast_2d(np.asarray(P, dtype=np.float64))
    return np.linalg.cholesky(0.5 * (P + P.T))


def srcubature(m, P):
    """Spherical-radial cubature: m +/- sqrt(d+1) L e_j (weight 1/(2(d+1))) and the centre m (weight 1/(d+1)),
    enumerated +e_1..+e_d, -e_1..-e_d, centre (ReactiveMP ``SphericalRadialCubature``)."""
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    d = m.size
    L = cholsqrt(np.reshape(P, (d, d)))
    s = np.sqrt(d + 1.0)
    pts = np.empty((2 * d + 1, d))
    for j in range(d):
        pts[j] = m + s * L[:, j]
        pts[d + j] = m - s * L[:, j]
    pts[2 * d] = m
    w = np.full(2 * d + 1, 1.0 / (2.0 * (d + 1)))
    w[2 * d] = 1.0 / (d + 1)
    return pts, w


def ghcubature(p, m, P):
    """Gauss-Hermite cubature of order p (ReactiveMP ``GaussHermiteCubature``): univariate points
    m + sqrt(2 v) t_i with weights w_i / sqrt(pi); multivariate = tensor grid through cholsqrt(P)."""
    t, w = np.polynomial.hermite.hermgauss(p)
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    d = m.size
    if d == 1:
        v = float(np.reshape(P, ()))
        return (m[0] + np.sqrt(2.0) * np.sqrt(v) * t)[:, None], w / np.sqrt(np.pi)
    L = cholsqrt(np.reshape(P, (d, d)))
    grids = np.meshgrid(*([np.arange(p)] * d), indexing="ij")
    idx = np.stack([g.ravel() for g in grids], axis=1)  # S x d
    pts = m[None, :] + np.sqrt(2.0) * (t[idx] @ L.T)
    wts = np.prod(w[idx], axis=1) / np.pi ** (d / 2.0)
    return pts, wts


def gen_unscented_uni(m, V, S=0.0, K=3.0):
    """helper_functions/ut_approx.jl:116-126 as written (S = skewness, K = kurtosis(q,false) = 3 for a Normal;
    the points land at m -/+ sqrt(3)/sqrt(V) -- dimensionally odd unless V = 1, reproduced as is)."""
    L = np.sqrt(V)
    invL3 = 1.0 / L**3
    u = 0.5 * (-S * invL3 + (1.0 / V) * np.sqrt(4.0 * K - 3.0 * S * S / V))
    v = u + S * invL3
    aux = 1.0 / (v * (u + v))
    pts = np.array([m, m - u * L, m + v * L], dtype=np.float64)[:, None]
    wts = np.array([1.0 - aux * (v / u + 1.0), (v / u) * aux, aux], dtype=np.float64)
    return pts, wts


def gen_unscented_multi(m, V, S=None, K=None):
    """helper_functions/ut_approx.jl:129-151.  ``cholinv(L.^3)`` is applied to a lower-triangular, non-symmetric
    matrix; FastCholesky factorises from the upper triangle, so only diag(L)^3 takes part (parity unpinned:
    FastCholesky.jl is not vendored).  Never reached by a live notebook."""
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    d = m.size
    S = np.zeros(d) if S is None else np.asarray(S, dtype=np.float64)
    K = np.full(d, 3.0) if K is None else np.asarray(K, dtype=np.float64)
    L = cholsqrt(np.reshape(V, (d, d)))
    l3 = np.diag(L) ** 3
    invL3 = 1.0 / l3
    invL4 = 1.0 / (l3 * l3)
    det = 4.0 * invL4 * K - 3.0 * (invL3 * S) ** 2
    u = 0.5 * (-invL3 * S + np.sqrt(det))
    v = u + invL3 * S
    pts = np.empty((2 * d + 1, d))
    wts = np.empty(2 * d + 1)
    pts[0] = m
    for i in range(d):
        pts[1 + i] = m - L[:, i] * u[i]
        pts[1 + d + i] = m + L[:, i] * v[i]
    wts[1 + d:] = 1.0 / v / (u + v)
    wts[1:1 + d] = wts[1 + d:] * (v / u)
    wts[0] = 1.0 - np.sum(wts[1:])
    return pts, wts


SRCUBATURE, GENUT, GAUSSHERMITE, CLOSED_FORM = 0, 1, 2, 3


def sigma_points(method, m, P, p=21):
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    if method == SRCUBATURE:
        return srcubature(m, P)
    if method == GAUSSHERMITE:
        return ghcubature(p, m, P)
    if method == GENUT:
        if m.size == 1:
            return gen_unscented_uni(float(m[0]), float(np.reshape(P, ())))
        return gen_unscented_multi(m, P)
    raise ValueError("no sigma points for method %r" % (method,))
