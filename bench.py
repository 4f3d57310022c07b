#!/usr/bin/env python
"""bench.py -- VSGP data-sweep throughput on B200 (contract in the task statement).

One "step" = one pass of the hot path over one batch: the generate-once K_uf + DMMA-SYRK sweep producing Psi0/Psi1/Psi2
for the kin40k-shape workload (N = 10000 points per GPU, D = 8, M = 512, SE-ARD, Float64).  `value` = data-points/s with
inputs resident in HBM (K steps enqueued back to back, an L2 flush on the stream before each, one CUDA event pair per step);
`e2e` = the same through the C-ABI call with HOST buffers (sgp_sweep_psi_host_packed: H2D of X / y -- beside the kernel's launch -- and D2H of
one packed buffer [lower triangle of Psi2 | Psi1 | scalars] inside the timed region; `e2e.full_square` = sgp_sweep_psi_host, Psi2 as the full square).  With --gpus N every rank sweeps its own N-shard and the statistics are summed over the ranks inside the
sweep kernel through NVLink peer memory ("weak": per-GPU work fixed); after the timed loop every rank checks its all-reduced
statistics against a single-GPU sweep of ALL ranks' points (`parity`, exit code 3 above 1e-12).  The synthetic N = 10M /
M = 1024 configuration, where the path is throughput-bound and the FP64 roofline is meaningful, is timed as well (strong scaling
over ranks) and reported under "synthetic_10M"; `dense` carries the M x M factorisations (N-th `prod`) beside LAPACK on the host.

--impl reference times the CPU port of the reference's per-point schedule (oracle/sweep_port.c; Julia is not installed
and the reference cannot be built here) on the host cores, on the same 10000-point workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KIN = dict(N=10000, D=8, M=512, ell=2.0, variance=1.0)
SYN = dict(N=10_000_000, D=8, M=1024, ell=2.0, variance=1.0)
FP64_SPEC_TFLOPS = 37.0   # HGX B200 FP64 / FP64-tensor spec ("40" on DGX B200)
PARITY_TOL = 1e-12


def fp64_peak_file():
    """FP64 DMMA peak measured on this pool's B200 in round 1 (tools/fp64_microbench.cu -> profiles/r01_fp64_peak.json)."""
    p = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["dmma_tflops"]), float(d["cublas_dgemm_tflops"])
    return FP64_SPEC_TFLOPS, None


def traffic(which):
    """(bytes per launch, source) of dram__bytes_read.sum + dram__bytes_write.sum of the sweep kernel from the newest committed
    `ncu --set full` capture (profiles/r02z_traffic.json, else earlier ones).  A static figure from a capture, NOT sampled in this run."""
    for name in ("r02z_traffic.json", "r02_traffic.json", "r01g_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        d = json.load(open(p))
        if which == "kin40k":
            return d.get("kin40k_bytes_per_launch"), "ncu capture profiles/%s: %s" % (name, d.get("kin40k_capture", ""))
        per_pt = d.get("synthetic_bytes_per_point")
        return (None if per_pt is None else per_pt * SYN["N"]), ("ncu capture profiles/%s, scaled per point to N=10M from: %s" % (name, d.get("synthetic_capture", "")))
    return None, "no capture committed"


class ClockSampler:
    def __init__(self, dev):
        self.dev = dev; self.samples = []; self.reasons = set(); self.stop = False; self.max_mhz = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.t.start(); return self

    def __exit__(self, *a):
        self.stop = True; self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def synth(cfg, n, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, cfg["D"]))
    beta = np.random.default_rng(2).standard_normal(cfg["D"])
    y = np.sin(X @ beta) + 0.1 * rng.standard_normal(n)
    return X, y


def inducing(cfg, M=None):
    # Z = rows of an X drawn with the data seed; chosen with seed 1
    M = M or cfg["M"]
    X0 = np.random.default_rng(0).standard_normal((max(4 * M, 4096), cfg["D"]))
    return X0[np.random.default_rng(1).choice(X0.shape[0], M, replace=False)].copy()


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def port_best_threads(cfg, Z, ell):
    """The C port of the reference schedule with 1, 4, 8, 16, ... , all host threads on a 300-point probe: the fastest setting."""
    from oracle import port
    X, y = synth(cfg, 300, 5)
    cands = sorted({1, 4, 8, 16, 32, host_threads()} & set(range(1, host_threads() + 1)) | {1})
    best = (0.0, 1)
    for th in cands:
        port.sweep(X[:50], y[:50], Z, cfg["variance"], ell, 1.0e4, threads=th)
        t0 = time.perf_counter()
        port.sweep(X, y, Z, cfg["variance"], ell, 1.0e4, threads=th)
        r = 300 / (time.perf_counter() - t0)
        if r > best[0]:
            best = (r, th)
    return best[1], best[0]


def run_reference(args):
    """The reference's own schedule (per-point rule + prod, GPnode/UniSGPnode.jl:144-158 + :62-63) as the C port, on the host cores, on the
    SAME kin40k-shape workload as the B200 arm: every step sweeps the 10000 points (bounded to fewer only if K + W steps would exceed ~4 min)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import port
    port.load(native=True)
    Z = inducing(KIN); ell = np.full(KIN["D"], KIN["ell"])
    threads, probe_rate = port_best_threads(KIN, Z, ell)
    budget_s = 240.0
    npts = KIN["N"]
    if (args.steps + args.warmup) * npts / probe_rate > budget_s:
        npts = max(500, int(budget_s * probe_rate / (args.steps + args.warmup)) // 500 * 500)
    X, y = synth(KIN, npts, 7)
    for _ in range(args.warmup):
        port.sweep(X, y, Z, KIN["variance"], ell, 1.0e4, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.sweep(X, y, Z, KIN["variance"], ell, 1.0e4, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = npts / dt
    sample = ("%d points per step (%s); C port of GPnode/UniSGPnode.jl:144-158 + :62-63 (Julia not installed), OpenMP over the kernel column, the rank-1 "
              "`mul!` and the M x M add of `prod`: %d thread(s), the fastest of a 1..%d-thread probe" %
              (npts, "the whole kin40k-shape batch: same config as the B200 arm" if npts == KIN["N"] else "bounded so that the run ends within minutes; per-point cost is constant",
               port.sweep.last_threads, host_threads()))
    line = {"impl": "reference", "metric": "vsgp_sweep_data_points_per_sec", "value": val, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "kin40k-shape VSGP sweep (BASELINE.json configs[1]): N=10000 points per GPU, D=8, M=512, SE-ARD, Float64; "
                                   "Psi0/Psi1/Psi2 per step", "same_config": npts == KIN["N"], "sample": "%d points per step" % npts},
            "cpu_baseline": {"value": val, "unit": "points/s", "cores": int(port.sweep.last_threads), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_best_effort(cfg, X, y, Z, ell):
    """Best-effort CPU: K_uf by a dgemm-form distance (|x|^2 + |z|^2 - 2 X Z'), in-place exp, Psi2 by dsyrk --
    no N x M x D temporaries; OpenBLAS with all its threads."""
    from scipy.linalg import blas
    t0 = time.perf_counter()
    Xs = X / ell; Zs = Z / ell
    K = blas.dgemm(-2.0, Xs, Zs, trans_b=True)                     # N x M
    K += np.einsum("nd,nd->n", Xs, Xs)[:, None]
    K += np.einsum("md,md->m", Zs, Zs)[None, :]
    np.multiply(K, -0.5, out=K); np.exp(K, out=K)
    if cfg["variance"] != 1.0:
        K *= cfg["variance"]
    psi2 = blas.dsyrk(1.0, K, trans=1, lower=1)                    # K' K, lower triangle
    psi1 = K.T @ y
    dt = time.perf_counter() - t0
    return dt, psi1, psi2


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return host_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-synthetic", action="store_true", help="skip the N=10M / M=1024 leg")
    ap.add_argument("--no-dense", action="store_true", help="skip the M x M leg")
    ap.add_argument("--syn-steps", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gaussianprocessnode_b200 import SGPContext

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_rows(a):
        """all ranks' rows of the host array `a`, concatenated in rank order (NCCL all_gather)"""
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return torch.cat(outs, 0).cpu().numpy()

    def same_bits_everywhere(*arrays):
        t = torch.from_numpy(np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in arrays])).cuda()
        mx = t.clone(); mn = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        return bool(torch.equal(mx, mn))

    def rel(a, b):
        return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))

    ctx = SGPContext(local)
    if world > 1:
        uid = [SGPContext.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])

    # ------------------------------------------------------------------ FP64 tensor-pipe peak of THIS device, measured in this run
    peak_file, dgemm_file = fp64_peak_file()
    try:
        peak_run = ctx.fp64_peak(150)
    except Exception:
        peak_run = None
    peak = peak_run if peak_run and peak_run > 1.0 else peak_file
    peak_src = ("DMMA.8x8x4 register-loop peak measured in this run on this device (sgp_fp64_peak): %.2f TFLOP/s; round-1 microbenchmark %.1f, cuBLAS DGEMM %s, "
                "spec %.1f (MEASURED_PEAKS.json has no FP64 entry)" % (peak, peak_file, dgemm_file, FP64_SPEC_TFLOPS))

    # ------------------------------------------------------------------ kin40k-shape leg (the headline metric)
    cfg = KIN
    X, y = synth(cfg, cfg["N"], 100 + rank)
    Z = inducing(cfg)
    ell = np.full(cfg["D"], cfg["ell"])
    ctx.set_kernel(cfg["variance"], ell); ctx.set_inducing(Z); ctx.set_data(X, y)
    ctx.sweep_timed_flushed(W, 256)
    barrier()
    with ClockSampler(local) as clk:
        t_wall0 = time.perf_counter()
        # K steps enqueued back to back on the library's stream: [256 MB L2 flush] [event] sweep (+ exchange over the ranks) [event];
        # the flush is outside the timed intervals, and no host synchronisation sits between the steps (ranks stay in lock step
        # through the exchange itself instead of accumulating host launch skew)
        # ... and nothing else inside a timed interval: the event pair around the main kernel alone (the roofline's launch duration) is taken in a
        # second pass of K steps (each event record costs about a microsecond of stream time)
        ms_step_loc, _ = ctx.sweep_timed_flushed(args.steps, 256, main_kernel=False)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        ms_step_inner_loc, ms_main_loc = ctx.sweep_timed_flushed(args.steps, 256)
        barrier()
    ms_step = max_over_ranks(ms_step_loc)
    ms_main = max_over_ranks(ms_main_loc)
    ms_step_inner = max_over_ranks(ms_step_inner_loc)
    info = ctx.last_sweep_info()
    value = world * cfg["N"] / (ms_step * 1e-3)

    # e2e through the C ABI with HOST buffers in pinned memory (sgp_pinned_alloc): per step the H2D copy of X / y
    # and the D2H read of Psi0 / Psi1 / Psi2 / sum_y2 (one sgp_sweep_psi_host call) are inside the timed region
    from gaussianprocessnode_b200 import pinned_empty
    Xp = pinned_empty(X.shape); Xp[...] = X
    yp = pinned_empty(y.shape); yp[...] = y
    # Psi2 comes back as its packed lower triangle (LAPACK 'L' packed storage, sgp_sweep_psi_host_packed): the running sum of `prod` is symmetric, so
    # that is all of it at half the bytes over the bus; the full-square form of the same call is timed beside it (`full_square`)
    Mb = cfg["M"]
    psi1p = pinned_empty((Mb,)); psi2p = pinned_empty((Mb, Mb), order="F"); statsk = pinned_empty((Mb * (Mb + 1) // 2 + Mb + 4,))
    n_e2e = max(10, min(args.steps, 50))

    def e2e_loop(outbuf, packed):
        for _ in range(3):
            ctx.sweep_psi_host(Xp, yp, out=outbuf, packed=packed)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            o = ctx.sweep_psi_host(Xp, yp, out=outbuf, packed=packed)     # H2D of X / y (beside the kernel's launch), sweep, D2H of the statistics
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) / n_e2e), o

    e2e_full_s, out_full = e2e_loop((psi1p, psi2p), False)
    e2e_s, out = e2e_loop(statsk, True)        # ONE buffer [packed lower triangle | Psi1 | scalars], ONE device-to-host copy
    from gaussianprocessnode_b200.sgp import pack_lower
    ref_stats = ctx.sweep_psi()            # the resident-data sweep the device-timed `value` runs
    e2e_same = bool(np.array_equal(out[2], pack_lower(ref_stats[2])) and np.array_equal(out[1], ref_stats[1]) and out[0] == ref_stats[0]
                    and np.array_equal(out_full[2], ref_stats[2]))
    h2d = X.nbytes + y.nbytes
    d2h = statsk.nbytes
    e2e_val = world * cfg["N"] / e2e_s

    # ------------------------------------------------------------------ parity of the sum over the ranks (world > 1)
    parity = None
    if world > 1:
        # >= 100 back-to-back exchanges first (monotonic epochs, rank skew), then the statistics every rank holds against ONE GPU sweeping
        # the concatenation of all ranks' points
        ctx.sweep_timed(100)
        p0, p1, p2, sy = ctx.sweep_psi()
        Xall = gather_rows(X); yall = gather_rows(y)
        single = SGPContext(local)
        single.set_kernel(cfg["variance"], ell); single.set_inducing(Z); single.set_data(Xall, yall)
        f0, f1, f2, fy = single.sweep_psi()
        parity = {"kin40k": {"psi2_rel_fro": rel(p2, f2), "psi1_rel_fro": rel(p1, f1), "psi0_rel": abs(p0 - f0) / abs(f0), "sum_y2_rel": abs(sy - fy) / abs(fy),
                             "bitwise_equal_across_ranks": same_bits_everywhere(p2, p1, [p0, sy]), "exchanges_before_check": 100 + args.steps + W + 2 * (n_e2e + 3) + 1,
                             "points_checked": int(Xall.shape[0])}}
        # synthetic shape (M = 1024): a 200k-point slice, sharded
        cs = SYN
        ns = 200_000 // world
        Xs_, ys_ = synth(cs, ns, 900 + rank); Zs_ = inducing(cs); ells_ = np.full(cs["D"], cs["ell"])
        ctx.set_kernel(cs["variance"], ells_); ctx.set_inducing(Zs_); ctx.set_data(Xs_, ys_)
        ctx.sweep_timed(5)
        q0, q1, q2, qy = ctx.sweep_psi()
        single.set_kernel(cs["variance"], ells_); single.set_inducing(Zs_); single.set_data(gather_rows(Xs_), gather_rows(ys_))
        g0, g1, g2, gy = single.sweep_psi()
        parity["synthetic_200k"] = {"psi2_rel_fro": rel(q2, g2), "psi1_rel_fro": rel(q1, g1), "psi0_rel": abs(q0 - g0) / abs(g0),
                                    "bitwise_equal_across_ranks": same_bits_everywhere(q2, q1, [q0, qy]), "points_checked": int(ns * world)}
        # uncertain inputs at the pendulum shape (N = 300 nodes, d = 2, M = 48, D_out = 2, srcubature): GPnode/MultiSGPnode.jl:290-328 summed over
        # the nodes of all ranks through the exchange kernel
        from gaussianprocessnode_b200 import SRCUBATURE
        rngp = np.random.default_rng(124)
        Npend, dpend, Mpend = 300, 2, 48
        means = rngp.normal(size=(Npend, dpend)); A_ = rngp.normal(size=(Npend, dpend, dpend)) * 0.1
        covs = A_ @ np.swapaxes(A_, 1, 2) + 1e-2 * np.eye(dpend); Rw = rngp.normal(size=(Npend, 2))
        gx, gy_ = np.meshgrid(np.linspace(-2.5, 2.5, 8), np.linspace(-2.5, 2.5, 6)); Zp = np.stack([gx.ravel(), gy_.ravel()], 1)
        lo = Npend * rank // world; hi = Npend * (rank + 1) // world
        ctx.set_kernel(1.0, np.array([1.0, 1.2])); ctx.set_inducing(Zp)
        u0, u1, u2, _ = ctx.sweep_psi_uncertain(SRCUBATURE, means[lo:hi], covs[lo:hi], R=Rw[lo:hi], D_out=2)
        single.set_kernel(1.0, np.array([1.0, 1.2])); single.set_inducing(Zp)
        v0, v1, v2, _ = single.sweep_psi_uncertain(SRCUBATURE, means, covs, R=Rw, D_out=2)
        parity["uncertain_pendulum"] = {"psi2_rel_fro": rel(u2, v2), "psi1_rel_fro": rel(u1, v1), "psi0_rel": abs(u0 - v0) / abs(v0),
                                        "bitwise_equal_across_ranks": same_bits_everywhere(u2, u1, [u0]), "nodes_checked": Npend}
        single.close()
        worst = max(max(v for k, v in d.items() if k.endswith("_rel_fro") or k.endswith("_rel")) for d in parity.values())
        parity["max_rel"] = worst
        parity["ok"] = bool(worst <= PARITY_TOL and all(d["bitwise_equal_across_ranks"] for d in parity.values() if isinstance(d, dict)))
        parity["tolerance"] = PARITY_TOL
        ctx.set_kernel(cfg["variance"], ell); ctx.set_inducing(Z); ctx.set_data(X, y)

    flops = cfg["N"] * cfg["M"] * (cfg["M"] + 1)
    achieved = flops / (ms_main * 1e-3) * 1e-12
    tr_k, tr_k_src = traffic("kin40k")
    line = {
        "metric": "vsgp_sweep_data_points_per_sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "kin40k-shape VSGP sweep (BASELINE.json configs[1]): N=10000 points per GPU, D=8, M=512, SE-ARD, Float64; "
                               "Psi0/Psi1/Psi2 per step" + ("; statistics summed over the ranks inside the sweep kernel (NVLink peer memory, two-shot all-reduce of the packed lower triangle); ranks aligned by a flag barrier after every L2 flush, before the timed interval" if world > 1 else ""),
                   "l2": "256 MB buffer rewritten on the stream before every timed step (inputs are smaller than L2); flush outside the timed intervals",
                   "parallelism": "N sharded over %d GPU(s)" % world, "wall_s_timed_region": t_wall},
        "e2e": {"value": e2e_val, "unit": "points/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_call": e2e_s * 1e3,
                "call": "sgp_sweep_psi_host_packed: pinned host X / y in (uploaded on the copy stream beside the kernel's launch), ONE packed buffer out -- [Psi2 as its lower triangle in LAPACK 'L' packed storage | Psi1 | Psi0, sum_y2, sum_w, n] with one device-to-host copy --, one host synchronisation",
                "bitwise_equal_to_resident_sweep": e2e_same,
                "full_square": {"value": world * cfg["N"] / e2e_full_s, "ms_per_call": e2e_full_s * 1e3, "d2h_bytes_per_step": int(out_full[2].nbytes + out_full[1].nbytes + 32),
                                "call": "sgp_sweep_psi_host (Psi2 as the full symmetric M x M square)"}},
        "gpu_launches": int(info["launches"] * args.steps),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": tr_k, "traffic_source": tr_k_src,
                     "frac_vs_round1_peak_%.1f" % peak_file: achieved / peak_file,
                     "kernel": "sweep4_kernel<128,32,8,256> (generate-once sweep: one cooperative launch per sweep) grid=%d block=%d smem=%d" % (info["grid"], info["block"], info["smem_bytes"]),
                     "algorithmic_flops_per_launch": flops, "ms_per_launch": ms_main, "peak_source": peak_src,
                     "ms_per_launch_source": "CUDA event pair around the kernel alone on the library's stream, mean over a second pass of %d flushed steps (ms_per_step of that pass, with the inner events: %.4f)" % (args.steps, ms_step_inner),
                     "note": "kin40k shape is 2.6 GFLOP: launch/latency-bound (71 us at peak); see synthetic_10M for the throughput-bound case"},
        "clocks": clk.summary(),
    }
    if parity is not None:
        line["parity"] = parity

    # ------------------------------------------------------------------ synthetic N = 10M, M = 1024 (strong scaling)
    if not args.no_synthetic:
        cfg = SYN
        n_loc = cfg["N"] // world
        g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
        n_pad = ((n_loc + 31) // 32) * 32
        Xd = torch.randn(n_pad, cfg["D"], dtype=torch.float64, device="cuda", generator=g)
        yd = torch.sin(Xd[:, 0] * 0.7 + Xd[:, 1] * 0.3) + 0.1 * torch.randn(n_pad, dtype=torch.float64, device="cuda", generator=g)
        Zs = inducing(cfg); ells = np.full(cfg["D"], cfg["ell"])
        ctx.set_kernel(cfg["variance"], ells); ctx.set_inducing(Zs)
        ctx.set_data_dev(n_pad, Xd.data_ptr(), yd.data_ptr())
        torch.cuda.synchronize()
        for _ in range(3):
            ctx.sweep_timed(1)
        barrier()
        with ClockSampler(local) as clk2:
            ts, tm = [], []
            for _ in range(args.syn_steps):
                a, b = ctx.sweep_timed(1)      # 640 MB of inputs per GPU at N=1: larger than L2
                ts.append(a); tm.append(b)
            barrier()
        ms_s = max_over_ranks(float(np.mean(ts))); ms_m = max_over_ranks(float(np.mean(tm)))
        fl = n_pad * cfg["M"] * (cfg["M"] + 1)
        info2 = ctx.last_sweep_info()
        tr_s, tr_s_src = traffic("synthetic")
        line["synthetic_10M"] = {
            "workload": "BASELINE.json configs[4]: N=10M (sharded: %d per GPU), D=8, M=1024, SE-ARD, Float64" % n_pad, "scaling": "strong",
            "value": world * n_pad / (ms_s * 1e-3), "unit": "points/s", "ms_per_step": ms_s, "steps": args.syn_steps,
            "psi2_tflops_all_gpus": world * fl / (ms_s * 1e-3) * 1e-12,
            "roofline": {"bound": "tensor", "achieved": fl / (ms_m * 1e-3) * 1e-12, "peak": peak, "unit": "TFLOP/s",
                         "frac": fl / (ms_m * 1e-3) * 1e-12 / peak, "traffic": tr_s, "traffic_source": tr_s_src, "ms_per_launch": ms_m,
                         "algorithmic_flops_per_launch": fl, "kernel": "sweep4_kernel<128,32,8,256> grid=%d" % info2["grid"],
                         "peak_source": peak_src},
            "clocks": clk2.summary()}
        # the same sweep with 32 MB panels (ring of 96 MB: fewer slab turnovers, but the K_uf panels no longer stay in L2 -- they are
        # written back to and partly re-read from HBM); reported beside the default, not instead of it
        # ... and with 16 MB panels (ring of 48 MB): the configuration whose DRAM traffic stays within 1.5x of the algorithmic bytes (ncu at N = 400k:
        # 36 + 18 MB per launch against 29 + 8; the default's 60 MB ring leaks ~6 % of its lines: 89 + 194 MB), about 2 % slower
        try:
            for mb, key, note in (("32", "slab_32MB_spilling", "K_uf panels round-trip HBM"), ("16", "slab_16MB_hbm_clean", "DRAM traffic <= 1.5x algorithmic (ncu, N = 400k: 36 + 18 MB per launch)")):
                os.environ["SGP_SWEEP_SLAB_MB"] = mb
                for _ in range(2):
                    ctx.sweep_timed(1)
                barrier()
                ts2 = [ctx.sweep_timed(1)[1] for _ in range(args.syn_steps)]
                ms2 = max_over_ranks(float(np.mean(ts2)))
                line["synthetic_10M"][key] = {"ms_per_launch": ms2, "tflops": fl / (ms2 * 1e-3) * 1e-12, "frac": fl / (ms2 * 1e-3) * 1e-12 / peak, "note": note}
        finally:
            os.environ.pop("SGP_SWEEP_SLAB_MB", None)
        del Xd, yd

    # ------------------------------------------------------------------ M x M leg: the N-th `prod` and K_uu (north_star part 4), N = 1 only
    if world == 1 and not args.no_dense:
        try:
            from scipy.linalg import lapack
            dense = {"what": "device ms per call (CUDA events on the library's stream around the call's kernels, resident inputs, mean of 10) and end to end "
                             "through the C ABI with pinned host buffers; LAPACK beside it on the host (scipy dpotrf + dpotri + dpotrf, %d BLAS threads)" % blas_threads(),
                     "sizes": {}}
            for Md in (512, 600, 1024):
                Xd_, yd_ = synth(KIN, 10000, 31); Zd = inducing(KIN, Md); elld = np.full(KIN["D"], KIN["ell"]); w = 1.0e4
                ctx.set_kernel(1.0, elld); ctx.set_inducing(Zd); ctx.set_data(Xd_, yd_)
                p0, p1, p2, sy = ctx.sweep_psi()
                ctx.prior_set_isotropic(50.0)
                ctx.kuu_factor(1e-8, fetch=False); ctx.posterior_v_stream(w, carry=False, fetch=True); ctx.w_terms(None, None)   # warm-up: allocations at this M
                t_kuu = ctx.dense_timed(0, jitter=1e-8); t_pu = ctx.dense_timed(1, w=w); t_p = ctx.dense_timed(2, w=w); t_w = ctx.dense_timed(3)
                Lp = pinned_empty((Md, Md), order="F"); Lp[...] = np.eye(Md) / 50.0
                outp = (pinned_empty((Md,)), pinned_empty((Md, Md), order="F"), pinned_empty((Md, Md), order="F"))
                xi0 = np.zeros(Md)
                for _ in range(2):
                    ctx.posterior_v(xi0, Lp, w, out=outp)
                t0 = time.perf_counter()
                for _ in range(10):
                    ctx.posterior_v(xi0, Lp, w, out=outp)
                t_e2e = (time.perf_counter() - t0) / 10 * 1e3
                Lam = np.asfortranarray(np.eye(Md) / 50.0 + w * p2)
                best = 1e30
                for _ in range(3):
                    t0 = time.perf_counter()
                    c, info_ = lapack.dpotrf(Lam, lower=1, clean=1)
                    Sg, _ = lapack.dpotri(c, lower=1)
                    Sg = np.tril(Sg) + np.tril(Sg, -1).T
                    mu_ = Sg @ (w * p1)
                    lapack.dpotrf(Sg + np.outer(mu_, mu_), lower=1, clean=1)
                    best = min(best, (time.perf_counter() - t0) * 1e3)
                dense["sizes"]["M=%d" % Md] = {"kuu_factor_ms": t_kuu, "posterior_v_with_Uv_ms": t_pu, "posterior_v_without_Uv_ms": t_p, "w_terms_ms": t_w,
                                               "posterior_v_e2e_host_buffers_ms": t_e2e, "cpu_lapack_posterior_ms": best,
                                               "h2d_bytes": int(8 * (Md * Md + Md)), "d2h_bytes": int(8 * (2 * Md * Md + Md))}
            line["dense"] = dense
        except Exception as e:  # pragma: no cover
            line["dense"] = {"error": repr(e)}

    # ------------------------------------------------------------------ theta step, N = 1 only
    if world == 1:
        try:
            cfg = KIN
            nb_ = 500                                   # the reference calls the gradient once per mini-batch of 500
            Xb, yb = synth(cfg, nb_, 11); Zb = inducing(cfg); ellb = np.full(cfg["D"], cfg["ell"])
            rngb = np.random.default_rng(5)
            vb = rngb.standard_normal(cfg["M"]); Cb = rngb.standard_normal((cfg["M"], cfg["M"])) * 0.05
            Uvb = np.linalg.cholesky(Cb @ Cb.T + 0.1 * np.eye(cfg["M"])).T
            ctx.set_kernel(cfg["variance"], ellb); ctx.set_inducing(Zb); ctx.set_data(Xb, yb)

            def th_step():
                ctx.set_kernel(cfg["variance"], ellb)   # theta changes every step in the driver: K_uu is refactored, the sweep re-run
                return ctx.theta_objective(vb, Uvb, 1.0e4, 1e-8)
            for _ in range(3):
                th_step()
            t0 = time.perf_counter()
            for _ in range(10):
                th_step()
            dt = (time.perf_counter() - t0) / 10
            from oracle import theta as otheta
            t0 = time.perf_counter(); otheta.neg_log_backwardmess_fast(cfg["variance"], ellb, yb, Xb, vb, Uvb, 1.0e4, Zb, 0, 1e-8)
            dt_cpu = time.perf_counter() - t0
            line["theta_step"] = {"workload": "objective + exact gradient of the theta step at a NEW theta, one kin40k mini-batch: N=500, D=8, M=512 (host buffers in, "
                                              "D+2 scalars out; sgp_theta_objective incl. the sweep and the K_uu factorisation)", "ms_per_call": dt * 1e3,
                                  "cpu_value_only_ms": dt_cpu * 1e3,
                                  "cpu_kind": "oracle restatement of neg_log_backwardmess_fast (value only, 1 thread; the reference adds ForwardDiff: "
                                              "3 chunked dual-number passes for 9 parameters)"}
            # one full mini-batch step of the kin40k driver (regression_kin40k.ipynb:196-228), device-resident streaming:
            # new theta -> upload 500 points -> sweep -> posterior on the resident prior (becomes the next prior) -> theta objective + gradient
            ctx.prior_set_isotropic(50.0)

            def mb_step():
                ctx.set_kernel(cfg["variance"], ellb)
                ctx.set_data(Xb, yb); ctx.sweep_psi(fetch=False)
                ctx.posterior_v_stream(1.0e4, carry=True)
                return ctx.theta_objective(None, None, 1.0e4, 0.0)
            for _ in range(3):
                mb_step()
            ctx.prior_set_isotropic(50.0)
            t0 = time.perf_counter()
            for _ in range(20):
                mb_step()
            torch.cuda.synchronize()
            dt_mb = (time.perf_counter() - t0) / 20
            line["minibatch_step"] = {"workload": "kin40k driver mini-batch (500 points, M=512) at a new theta: H2D of the batch, sweep, posterior on the resident "
                                                  "streaming prior, K_uu factorisation, theta objective + gradient; only D+2 scalars return to the host",
                                      "ms_per_minibatch": dt_mb * 1e3, "points_per_s": 500 / dt_mb,
                                      "reference": "~400 points/s end to end (3h30 for 500 epochs x 10000 points, regression_kin40k.ipynb:239; "
                                                   "author's Mac, includes RxInfer scheduling)"}
        except Exception as e:  # pragma: no cover
            line["theta_step"] = {"error": repr(e)}

    # ------------------------------------------------------------------ CPU baselines (rank 0, N = 1 only)
    if rank == 0 and world == 1:
        from oracle import port
        port.load(native=True)
        Zk = inducing(KIN); ellk = np.full(KIN["D"], KIN["ell"])
        threads, _ = port_best_threads(KIN, Zk, ellk)
        Xc, yc = synth(KIN, KIN["N"], 7)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter(); port.sweep(Xc, yc, Zk, KIN["variance"], ellk, 1.0e4, threads=threads); best = min(best, time.perf_counter() - t0)
        used = int(port.sweep.last_threads)
        t0 = time.perf_counter(); port.sweep(Xc[:2000], yc[:2000], Zk, KIN["variance"], ellk, 1.0e4, threads=1); r1 = 2000 / (time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": KIN["N"] / best, "unit": "points/s", "cores": used, "kind": "port",
                                "sample": "all 10000 kin40k-shape points, best of 3 passes (%.2f s each); C port of the reference's per-point rule + prod "
                                          "schedule (oracle/sweep_port.c), OpenMP over the column / rank-1 `mul!` / M x M add with %d thread(s) (fastest of a probe; "
                                          "1 thread: %.0f points/s); Julia not installed" % (best, used, r1)}
        try:
            dt, c1, c2 = cpu_best_effort(KIN, Xc, yc, Zk, ellk)
            dt = min(dt, cpu_best_effort(KIN, Xc, yc, Zk, ellk)[0])
            line["cpu_best_effort"] = {"value": KIN["N"] / dt, "unit": "points/s", "cores": blas_threads(),
                                       "kind": "NOT the reference's schedule: K_uf by dgemm-form distances + in-place exp, Psi2 by OpenBLAS dsyrk, Psi1 by dgemv "
                                               "(BASELINE.md section 3 item 2); %d BLAS threads" % blas_threads()}
        except Exception as e:  # pragma: no cover
            line["cpu_best_effort"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
