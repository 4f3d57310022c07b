"""The theta step of the kin40k driver (N=500, D=8, M=512 by default) in a loop: for a kernel launch list (ncu) and wall-clock per call.
usage: theta_profile.py [M] [N] [resident|host]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext

M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500
mode = sys.argv[3] if len(sys.argv) > 3 else "host"
D = 8
rng = np.random.default_rng(0)
X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]); Z = rng.normal(size=(M, D))
ell = np.full(D, 2.0)
ctx = SGPContext(0)
ctx.set_kernel(1.0, ell); ctx.set_inducing(Z); ctx.set_data(X, y)
ctx.sweep_psi(fetch=False); ctx.prior_set_isotropic(50.0); ctx.kuu_factor(1e-8, fetch=False)
mu, Sig, Uv = ctx.posterior_v_stream(1e4, carry=False, fetch=True)


def step():
    ctx.set_kernel(1.0, ell)
    return ctx.theta_objective(mu, Uv, 1e4, 1e-8) if mode == "host" else ctx.theta_objective(None, None, 1e4, 1e-8)


for _ in range(3):
    step()
reps = int(os.environ.get("REPS", "10"))
t0 = time.perf_counter()
for _ in range(reps):
    step()
print("theta step M=%d N=%d (%s posterior): %.3f ms per call" % (M, N, mode, (time.perf_counter() - t0) / reps * 1e3))
ctx.close()
