"""Timing of the M x M factorisation entry points through the C ABI (host buffers in/out).  usage: profile_dense.py [M...]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext

Ms = [int(a) for a in sys.argv[1:]] or [512, 1024]
rng = np.random.default_rng(0)
for M in Ms:
    D, N = 8, 10000
    X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]); Z = X[:M].copy()
    ctx = SGPContext(0)
    ctx.set_kernel(1.0, np.full(D, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
    ctx.sweep_psi(fetch=False)
    xi0 = np.zeros(M); Lam0 = np.eye(M) / 50.0

    def timeit(f, reps=10):
        for _ in range(2): f()
        t0 = time.perf_counter()
        for _ in range(reps): f()
        return (time.perf_counter() - t0) / reps * 1e3
    t_kuu = timeit(lambda: ctx.kuu_factor(1e-8, fetch=False))
    t_kuu_f = timeit(lambda: ctx.kuu_factor(1e-8))
    t_post = timeit(lambda: ctx.posterior_v(xi0, Lam0, 100.0))
    from gaussianprocessnode_b200 import pinned_empty
    Lp = pinned_empty((M, M), order="F"); Lp[...] = Lam0
    outp = (pinned_empty((M,)), pinned_empty((M, M), order="F"), pinned_empty((M, M), order="F"))
    t_post_pin = timeit(lambda: ctx.posterior_v(xi0, Lp, 100.0, out=outp))
    mu, Sig, Uv = ctx.posterior_v(xi0, Lam0, 100.0)
    t_w = timeit(lambda: ctx.w_terms(mu, Uv))
    t_th = timeit(lambda: ctx.theta_objective(mu, Uv, 100.0, 1e-8))
    t_thv = timeit(lambda: ctx.theta_objective(mu, Uv, 100.0, 1e-8, grad=False))
    print("M=%4d: kuu_factor %.3f ms (with L to host %.3f) | posterior_v %.3f (pinned buffers %.3f) | w_terms %.3f | theta value %.3f, value+grad %.3f (N=%d)"
          % (M, t_kuu, t_kuu_f, t_post, t_post_pin, t_w, t_thv, t_th, N))
    ctx.close()
