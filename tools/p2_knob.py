"""Phase-2 plan knob of the generate-once sweep at the kin40k bench shape: SGP_SWEEP4_P2_DIAG = reduction cost of a diagonal tile's stripe relative to an
off-diagonal one (decides how many CTAs each tile gets after the grid barrier).  Flushed timing as in bench.py (200 steps, L2 flush before each)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext

N, D, M = 10000, 8, 512
rng = np.random.default_rng(0)
X = rng.standard_normal((N, D)); y = np.sin(X[:, 0]); Z = X[:M].copy()
ctx = SGPContext(0)
ctx.set_kernel(1.0, np.full(D, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
ref = None
for rnd in range(2):
    for v in sys.argv[1:] or ["0.5", "0.6", "0.75", "0.9", "1.0", "1.25"]:
        os.environ["SGP_SWEEP4_P2_DIAG"] = v
        ctx.sweep_timed_flushed(20, 256, main_kernel=False)
        ms, _ = ctx.sweep_timed_flushed(200, 256, main_kernel=False)
        out = ctx.sweep_psi()
        if ref is None:
            ref = out
        same = np.array_equal(out[2], ref[2]) and np.array_equal(out[1], ref[1])
        print("P2_DIAG=%-5s %.4f ms per step   bits equal to the first plan's: %s" % (v, ms, same))
ctx.close()
