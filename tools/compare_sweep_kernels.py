"""Same-process timing of the two sweep kernels (SGP_SWEEP_IMPL=3: first fused kernel, 4: generate-once) over a list of shapes N,M,D."""
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from gaussianprocessnode_b200 import SGPContext
ctx = SGPContext(0)
shapes = [(50, 20, 1), (1500, 48, 2), (4000, 64, 2), (4000, 500, 2), (500, 512, 8), (2000, 512, 8), (10000, 600, 8), (100000, 256, 8), (30000, 1024, 8)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for (N, M, D) in shapes:
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, D)); y = np.sin(X[:, 0]); Z = rng.standard_normal((M, D))
    ctx.set_kernel(1.0, np.full(D, 2.0), D=D); ctx.set_inducing(Z); ctx.set_data(X, y)
    out = []
    for impl in ("3", "4"):
        os.environ["SGP_SWEEP_IMPL"] = impl
        ctx.sweep_timed(3)
        ms = min(ctx.sweep_timed(20)[0] for _ in range(3))
        out.append(ms)
    print("N=%6d M=%4d D=%d  first fused kernel %.4f ms   generate-once %.4f ms   ratio %.2f" % (N, M, D, out[0], out[1], out[0] / out[1]))
ctx.close()
