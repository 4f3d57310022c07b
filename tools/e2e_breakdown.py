"""Where the end-to-end time of sgp_sweep_psi_host goes (kin40k bench shape: N=10000, D=8, M=512): wall-clock of the pieces, each synchronised."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext, pinned_empty

N, D, M = 10000, 8, 512
rng = np.random.default_rng(0)
X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]); Z = rng.normal(size=(M, D))
ctx = SGPContext(0)
ctx.set_kernel(1.0, np.full(D, 2.0)); ctx.set_inducing(Z)
Xp = pinned_empty(X.shape); Xp[...] = X
yp = pinned_empty(y.shape); yp[...] = y
psi1p = pinned_empty((M,)); psi2p = pinned_empty((M, M), order="F")


def timeit(f, reps=50):
    for _ in range(5): f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e3


psi2k = pinned_empty((M * (M + 1) // 2 + M + 4,))
print("sweep_psi_host (H2D beside the launch + sweep + D2H)   %.3f ms" % timeit(lambda: ctx.sweep_psi_host(Xp, yp, out=(psi1p, psi2p))))
print("sweep_psi_host packed Psi2                              %.3f ms" % timeit(lambda: ctx.sweep_psi_host(Xp, yp, out=psi2k, packed=True)))
os.environ["SGP_HOST_OVERLAP"] = "0"
print("sweep_psi_host, upload then launch (SGP_HOST_OVERLAP=0) %.3f ms" % timeit(lambda: ctx.sweep_psi_host(Xp, yp, out=(psi1p, psi2p))))
print("... packed                                              %.3f ms" % timeit(lambda: ctx.sweep_psi_host(Xp, yp, out=psi2k, packed=True)))
del os.environ["SGP_HOST_OVERLAP"]
print("set_data (H2D + sync)               %.3f ms" % timeit(lambda: ctx.set_data(Xp, yp)))
print("sweep_psi(fetch=False) (sweep, sync) %.3f ms" % timeit(lambda: ctx.sweep_psi(fetch=False)))
print("sweep_psi (sweep + D2H)             %.3f ms" % timeit(lambda: ctx.sweep_psi(out=(psi1p, psi2p))))
print("device time of the sweep            %.3f ms" % ctx.sweep_timed(20)[0])
ctx.close()
