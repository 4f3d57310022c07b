import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
from gaussianprocessnode_b200 import SGPContext, pinned_empty
ctx = SGPContext(0)
rng = np.random.default_rng(0)
N, D, M = 10000, 8, 512
X = pinned_empty((N, D)); X[...] = rng.standard_normal((N, D)); y = pinned_empty((N,)); y[...] = np.sin(X[:, 0])
Z = rng.standard_normal((M, D))
psi1 = pinned_empty((M,)); psi2 = pinned_empty((M, M), order="F")
ctx.set_kernel(1.0, np.full(D, 2.0)); ctx.set_inducing(Z)
for _ in range(5): ctx.sweep_psi_host(X, y, out=(psi1, psi2))
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    for _ in range(200): ctx.sweep_psi_host(X, y, out=(psi1, psi2))
    best = min(best, (time.perf_counter() - t0) / 200)
print("e2e %.4f ms per call  (%.1f M points/s)" % (best * 1e3, N / best * 1e-6))
ctx.close()
