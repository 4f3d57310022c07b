"""Timing of the uncertain-input statistics call (sgp_sweep_psi_uncertain) through the C ABI with host buffers.
usage: profile_uncertain.py [N] [M] [d] [reps]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext, SRCUBATURE, GENUT, GAUSSHERMITE, CLOSED_FORM_SE

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
M = int(sys.argv[2]) if len(sys.argv) > 2 else 48
d = int(sys.argv[3]) if len(sys.argv) > 3 else 2
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
rng = np.random.default_rng(0)
mean = rng.normal(size=(N, d)) * 1.5
A = rng.normal(size=(N, d, d)) * 0.1
cov = A @ np.swapaxes(A, 1, 2) + 1e-3 * np.eye(d)
Z = rng.normal(size=(M, d)) * 2.0
Y = rng.normal(size=(N, 2))
ctx = SGPContext(0)
ctx.set_kernel(1.3, np.full(d, 1.2)); ctx.set_inducing(Z)
for name, method, kw in (("srcubature D_out=2", SRCUBATURE, dict(R=Y, D_out=2)), ("srcubature D_out=1", SRCUBATURE, {}), ("genut", GENUT, {}),
                         ("gauss-hermite p=5", GAUSSHERMITE, dict(p=5)), ("closed form", CLOSED_FORM_SE, {})):
    for _ in range(3):
        ctx.sweep_psi_uncertain(method, mean, cov, **kw)
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.sweep_psi_uncertain(method, mean, cov, **kw)
    dt = (time.perf_counter() - t0) / reps
    print("N=%d M=%d d=%d %-20s %.3f ms per call  (%.2f M inputs/s)" % (N, M, d, name, dt * 1e3, N / dt * 1e-6))
ctx.close()
