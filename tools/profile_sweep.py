"""Small driver for ncu: the synthetic M=1024 / D=8 sweep on N points (default 1M), a few launches."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8)); y = np.sin(X[:, 0])
Z = X[np.random.default_rng(1).choice(N, M, replace=False)].copy()
ctx = SGPContext(0)
ctx.set_kernel(1.0, np.full(8, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
for _ in range(reps):
    ms, ms_main = ctx.sweep_timed(1)
    fl = N * M * (M + 1)
    print("N=%d M=%d sweep %.3f ms main %.3f ms  %.2f TFLOP/s  %s" % (N, M, ms, ms_main, fl / ms_main * 1e-9, ctx.last_sweep_info()))
ctx.close()
