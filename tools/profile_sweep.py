"""Small driver for timing / ncu: the synthetic D=8 sweep on N points (default 1M, M=1024), a few launches.
usage: profile_sweep.py [N] [M] [reps] [--clocks] [--check]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessnode_b200 import SGPContext

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if len(args) > 0 else 1_000_000
M = int(args[1]) if len(args) > 1 else 1024
reps = int(args[2]) if len(args) > 2 else 3
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8)); y = np.sin(X[:, 0])
Z = X[np.random.default_rng(1).choice(N, M, replace=False)].copy()
ctx = SGPContext(0)
ctx.set_kernel(1.0, np.full(8, 2.0)); ctx.set_inducing(Z); ctx.set_data(X, y)
if "--clocks" in sys.argv:
    ctx.sweep_debug_clocks()
for _ in range(reps):
    ms, ms_main = ctx.sweep_timed(1)
    fl = N * M * (M + 1)
    print("N=%d M=%d sweep %.3f ms main %.3f ms  %.2f TFLOP/s  %s" % (N, M, ms, ms_main, fl / ms_main * 1e-9, ctx.last_sweep_info()))
if "--clocks" in sys.argv:
    r = ctx.sweep_debug_clocks()
    r = r[r[:, 0] >= 0]
    if os.environ.get("SGP_SWEEP_IMPL", "4") != "3":
        r = r.astype(float); r[r[:, 2] < 2, 0] /= 8.0        # the generate-once kernel records k-steps (1/8 chunk)
    for dg in (0, 1):
        q = r[r[:, 2] == dg]
        if len(q):
            print("diag=%d: %d segments, clocks/chunk mean %.0f (min %.0f max %.0f)" % (dg, len(q), (q[:, 1].sum() / q[:, 0].sum()),
                  (q[:, 1] / q[:, 0]).min(), (q[:, 1] / q[:, 0]).max()))
    for dg, name in ((2, "generator (clocks per generated block-chunk)"), (3, "dependency waits (clocks per sweep)")):
        q = r[r[:, 2] == dg]
        if len(q):
            print("%s: mean %.0f max %.0f; per-CTA total clocks mean %.3e max %.3e" % (name, q[:, 1].sum() / max(q[:, 0].sum(), 1),
                  (q[:, 1] / np.maximum(q[:, 0], 1)).max(), q[:, 1].mean(), q[:, 1].max()))
    q4, q5 = r[r[:, 2] == 4], r[r[:, 2] == 5]
    if len(q4):
        print("timeline (clocks, mean / max over CTAs): setup %.0f / %.0f, slab loop %.0f / %.0f, final barrier %.0f / %.0f, phase 2 %.0f / %.0f" % (
              q4[:, 0].mean(), q4[:, 0].max(), q4[:, 1].mean(), q4[:, 1].max(), q5[:, 0].mean(), q5[:, 0].max(), q5[:, 1].mean(), q5[:, 1].max()))
    q6 = r[r[:, 2] == 6]
    if len(q6):
        print("phase 2 (clocks, mean / max over CTAs): slot discovery %.0f / %.0f, partial loads %.0f / %.0f, stores %.0f / %.0f" % (
              q6[:, 0].mean(), q6[:, 0].max(), q6[:, 1].mean(), q6[:, 1].max(), q6[:, 3].mean(), q6[:, 3].max()))
    q7 = r[r[:, 2] == 7]
    if len(q7):
        print("slab loop besides the generator (clocks per sweep, mean over CTAs): first-segment prologue %.0f, segments %.0f, closing barrier %.0f" % (
              q7[:, 0].mean(), q7[:, 1].mean(), q7[:, 3].mean()))
    r = r[r[:, 2] < 2]
    for dg in (0, 1):       # least-squares  clocks = a * chunks + b  over the segments (per slab pass: b is per segment and slab)
        q = r[r[:, 2] == dg].astype(float)
        if len(q) > 2 and q[:, 0].std() > 0:
            A = np.stack([q[:, 0], np.ones(len(q))], 1)
            (a_, b_), *_ = np.linalg.lstsq(A, q[:, 1], rcond=None)
            print("diag=%d fit: %.0f clocks per chunk + %.0f per segment (summed over slabs)" % (dg, a_, b_))
    if "--dump" in sys.argv:
        os.makedirs("gpurun_out", exist_ok=True)
        np.savetxt("gpurun_out/segments_N%d_M%d.csv" % (N, M), r, fmt="%d", delimiter=",", header="chunks,clocks,diag,cta")
    per_cta = {}
    for ch, clk, dg, cta in r:
        per_cta[cta] = per_cta.get(cta, 0) + clk
    v = np.array(list(per_cta.values()), dtype=float)
    print("per-CTA clocks: min %.3e mean %.3e max %.3e (%d CTAs)" % (v.min(), v.mean(), v.max(), len(v)))
if "--check" in sys.argv:
    n = min(N, 20000)
    ctx.set_data(X[:n], y[:n])
    psi0, psi1, psi2, sy2 = ctx.sweep_psi()
    Xs, Zs = X[:n] / 2.0, Z / 2.0
    d2 = (Xs ** 2).sum(1)[None, :] + (Zs ** 2).sum(1)[:, None] - 2 * Zs @ Xs.T
    K = np.exp(-0.5 * np.maximum(d2, 0))
    print("check n=%d: psi2 relF %.2e  psi1 rel %.2e" % (n, np.linalg.norm(psi2 - K @ K.T) / np.linalg.norm(K @ K.T),
          np.linalg.norm(psi1.ravel() - K @ y[:n]) / np.linalg.norm(K @ y[:n])))
ctx.close()
