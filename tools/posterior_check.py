"""One posterior of v at M inducing points through the C ABI; prints the residual of Uv'Uv = Sigma + mu mu' and cond(Sigma).
usage: [SGP_DENSE_UV2=1] posterior_check.py M   (debugging aid for the one-launch / two-launch forms of sgp_posterior_v)"""
import os
import sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext
M = int(sys.argv[1]); rng = np.random.default_rng(7 + M)
X = rng.normal(size=(4000, 4)); y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=4000)
Z = X[rng.choice(4000, M, replace=False)]
c = SGPContext(0); c.set_kernel(1.1, np.full(4, 1.4), D=4); c.set_inducing(Z); c.set_data(X, y); c.sweep_psi()
mu, Sig, Uv = c.posterior_v(np.zeros(M), np.eye(M) / 50.0, 40.0)
print("ok", np.linalg.norm(Uv.T @ Uv - Sig - np.outer(mu, mu)) / np.linalg.norm(Sig), np.linalg.cond(Sig))
