"""Per-phase clocks of the M x M job kernel (SGP_DENSE_CLOCKS=1 makes libsgp print CTA 0's counters).  usage: dense_clocks.py [M...]"""
import os, sys
os.environ["SGP_DENSE_CLOCKS"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext

rng = np.random.default_rng(0)
for M in [int(a) for a in sys.argv[1:]] or [512, 1024]:
    D, N = 8, 10000
    X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]); Z = X[:M].copy()
    ctx = SGPContext(0)
    ctx.set_kernel(1.0, np.full(D, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
    ctx.sweep_psi(fetch=False)
    ctx.prior_set_isotropic(50.0)
    for _ in range(2):
        ctx.kuu_factor(1e-8, fetch=False)
        ctx.posterior_v_stream(1.0e4, carry=False, fetch=True)
    ctx.close()
