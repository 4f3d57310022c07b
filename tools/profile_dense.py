"""Timing of the M x M entry points: device time per call (sgp_dense_timed) and end to end through the C ABI (host buffers in / out).
usage: profile_dense.py [M...]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext, pinned_empty

Ms = [int(a) for a in sys.argv[1:]] or [512, 1024]
rng = np.random.default_rng(0)
for M in Ms:
    D, N = 8, 10000
    X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]); Z = X[:M].copy()
    ctx = SGPContext(0)
    ctx.set_kernel(1.0, np.full(D, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
    ctx.sweep_psi(fetch=False)
    ctx.prior_set_isotropic(50.0)
    ctx.kuu_factor(1e-8, fetch=False); ctx.posterior_v_stream(1e4, carry=False, fetch=True)      # warm-up: allocations, lazy module load
    names = ["kuu_factor", "posterior with Uv", "posterior without Uv", "w_terms"]
    for what in range(4):
        t = [ctx.dense_timed(what, w=1.0e4, jitter=1e-8, reps=1) for _ in range(8)]
        print("M=%4d %-22s device ms per call: %s" % (M, names[what], " ".join("%.3f" % v for v in t)))
    xi0 = np.zeros(M); Lp = pinned_empty((M, M), order="F"); Lp[...] = np.eye(M) / 50.0
    outp = (pinned_empty((M,)), pinned_empty((M, M), order="F"), pinned_empty((M, M), order="F"))

    def timeit(f, reps=10):
        for _ in range(2): f()
        t0 = time.perf_counter()
        for _ in range(reps): f()
        return (time.perf_counter() - t0) / reps * 1e3
    print("M=%4d end to end (pinned host buffers): posterior_v %.3f ms | posterior_v without Uv %.3f ms | kuu_factor (no fetch) %.3f ms | theta value+grad %.3f ms" % (
        M, timeit(lambda: ctx.posterior_v(xi0, Lp, 1e4, out=outp)), timeit(lambda: ctx.posterior_v(xi0, Lp, 1e4, want_Uv=False, out=(outp[0], outp[1], None))),
        timeit(lambda: ctx.kuu_factor(1e-8, fetch=False)), timeit(lambda: ctx.theta_objective(None, None, 1e4, 1e-8))))
    ctx.close()
