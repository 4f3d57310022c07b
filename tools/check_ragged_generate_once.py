import os, sys, numpy as np
os.environ["SGP_SWEEP_IMPL"] = "4"
sys.path.insert(0, os.getcwd())
from gaussianprocessnode_b200 import SGPContext
from oracle import batched
ctx = SGPContext(0)
worst = 0.0
for (N, D, M) in [(1, 1, 1), (31, 3, 7), (33, 2, 65), (1000, 8, 129), (4097, 5, 200), (777, 12, 300), (50, 16, 48), (5, 8, 513), (100, 1, 1025), (70000, 16, 700), (3000, 4, 1500)]:
    rng = np.random.default_rng(N + D + M)
    X = rng.normal(size=(N, D)); Z = rng.normal(size=(M, D)); y = rng.normal(size=N); w = rng.uniform(0.5, 2, N); yv = rng.uniform(0, 1, N)
    ell = 0.7 + rng.random(D) * 2.0
    for wts in (None, w):
        ctx.set_kernel(1.7, ell, D=D); ctx.set_inducing(Z); ctx.set_data(X, y, yv, wts)
        p0, p1, p2, sy = ctx.sweep_psi()
        o0, o1, o2, oy = batched.psi_stats_point(X, y, Z, 1.7, ell, weights=wts, yvar=yv)
        e = max(np.linalg.norm(p2 - o2) / max(np.linalg.norm(o2), 1e-300), np.linalg.norm(p1 - o1) / max(np.linalg.norm(o1), 1e-300), abs(p0 - o0) / abs(o0), abs(sy - oy) / abs(oy))
        worst = max(worst, e)
        print(N, D, M, wts is not None, "%.2e" % e, ctx.last_sweep_info()["smem_bytes"])
print("worst", worst)
assert worst < 1e-10
