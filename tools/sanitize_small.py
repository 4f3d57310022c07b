"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): tiny shapes so that
the instrumented run finishes in seconds.  usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext, SRCUBATURE, CLOSED_FORM_SE, MATERN52

rng = np.random.default_rng(0)
ctx = SGPContext(0)
for (N, D, M, kind) in ((700, 8, 300, 0), (130, 3, 70, MATERN52)):
    X = rng.normal(size=(N, D)); y = rng.normal(size=N); Z = rng.normal(size=(M, D)); w = rng.uniform(0.5, 1.5, N)
    ctx.set_kernel(1.0, np.full(D, 1.5), kind=kind); ctx.set_inducing(Z)
    ctx.set_data(X, y, None, w); ctx.sweep_psi()
    ctx.set_data(X, y); p = ctx.sweep_psi()
    L = ctx.kuu_factor(1e-6)
    mu, Sig, Uv = ctx.posterior_v(np.zeros(M), np.eye(M) / 10.0, 5.0)
    ctx.w_terms(mu, Uv)
    ctx.theta_objective(mu, Uv, 5.0, 1e-6)
    ctx.kuu_solve(np.eye(M)[:, :3])
    ctx.predict_mean(X[:50], mu)
# the generate-once sweep kernel (default for M > 384; forced here on small shapes too, with tiny panels -> many slabs, ring wrap-around)
os.environ["SGP_SWEEP_IMPL"] = "4"; os.environ["SGP_SWEEP_SLAB_MB"] = "0.2"
for (N, D, M, kind) in ((900, 8, 520, 0), (400, 3, 70, MATERN52), (300, 12, 200, 0)):
    X = rng.normal(size=(N, D)); y = rng.normal(size=N); Z = rng.normal(size=(M, D)); w = rng.uniform(0.5, 1.5, N)
    ctx.set_kernel(1.0, np.full(D, 1.5), kind=kind); ctx.set_inducing(Z)
    ctx.set_data(X, y, None, w); ctx.sweep_psi()
    ctx.set_data(X, y); ctx.sweep_psi()
del os.environ["SGP_SWEEP_IMPL"]; del os.environ["SGP_SWEEP_SLAB_MB"]
d = 2
mean = rng.normal(size=(60, d)); A = rng.normal(size=(60, d, d)) * 0.1; cov = A @ np.swapaxes(A, 1, 2) + 1e-3 * np.eye(d)
ctx.set_kernel(1.0, np.full(d, 1.2)); ctx.set_inducing(rng.normal(size=(40, d)))
ctx.sweep_psi_uncertain(SRCUBATURE, mean, cov, R=rng.normal(size=(60, 2)), D_out=2, want_psi1_n=True)
ctx.kuu_factor(1e-8)
M2 = 40
ctx.uncertain_node_terms(rng.normal(size=M2), np.triu(rng.normal(size=(M2, M2))), 60)
ctx.in_logmessage(rng.normal(size=(60, 5, d)), rng.normal(size=(M2, 2)), np.eye(M2), 3.0, R=rng.normal(size=(60, 2)), hess=True)
ctx.sweep_psi_uncertain(CLOSED_FORM_SE, mean, cov)
ctx.close()
print("sanitize_small: done")
