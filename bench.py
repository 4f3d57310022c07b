#!/usr/bin/env python
"""bench.py -- VSGP data-sweep throughput on B200 (contract in the task statement).

One "step" = one pass of the hot path over one batch: the generate-once K_uf + DMMA-SYRK sweep producing Psi0/Psi1/Psi2
for the kin40k-shape workload (N = 10000 points per GPU, D = 8, M = 512, SE-ARD, Float64).  `value` = data-points/s with
inputs resident in HBM (K steps enqueued back to back, an L2 flush on the stream before each, one CUDA event pair per step);
`e2e` = the same through the C-ABI call with HOST buffers (sgp_sweep_psi_host: H2D of X / y and D2H of Psi1 / Psi2 inside
the timed region).  With --gpus N every rank sweeps its own N-shard and the statistics are summed over the ranks inside the
sweep kernel through NVLink peer memory ("weak": per-GPU work fixed).  The synthetic N = 10M / M = 1024
configuration, where the path is throughput-bound and the FP64 roofline is meaningful, is timed as well (strong scaling
over ranks) and reported under "synthetic_10M".

--impl reference times the CPU port of the reference's per-point schedule (oracle/sweep_port.c; Julia is not installed
and the reference cannot be built here) on the host cores.
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


def fp64_peak():
    """FP64 DMMA peak measured on this pool's B200 (tools/fp64_microbench.cu -> profiles/r01_fp64_peak.json).
    MEASURED_PEAKS.json has no FP64 entry (bf16 and HBM only)."""
    p = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["dmma_tflops"]), "measured DMMA peak (profiles/r01_fp64_peak.json; cuBLAS DGEMM %.1f)" % d["cublas_dgemm_tflops"]
    return FP64_SPEC_TFLOPS, "spec (no measurement file)"


def traffic(which):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of sweep_kernel from the committed `ncu --set full` captures
    (profiles/r01g_traffic.json: kin40k shape as benchmarked; synthetic scaled per point from the N=400000 capture)."""
    p = os.path.join(ROOT, "profiles", "r01g_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if which == "kin40k":
        return d.get("kin40k_bytes_per_launch")
    per_pt = d.get("synthetic_bytes_per_point")
    return None if per_pt is None else per_pt * SYN["N"]


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


def inducing(cfg):
    # Z = rows of an X drawn with the data seed; chosen with seed 1
    X0 = np.random.default_rng(0).standard_normal((max(4 * cfg["M"], 4096), cfg["D"]))
    return X0[np.random.default_rng(1).choice(X0.shape[0], cfg["M"], replace=False)].copy()


def cpu_port_rate(cfg, npts, native=True, reps=1):
    from oracle import port
    X, y = synth(cfg, npts, 7)
    Z = inducing(cfg)
    ell = np.full(cfg["D"], cfg["ell"])
    port.load(native=native)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        port.sweep(X, y, Z, cfg["variance"], ell, 1.0e4, native=False)
        best = min(best, time.perf_counter() - t0)
    return npts / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    npts = 500                                  # one reference mini-batch (regression_kin40k.ipynb: batch_size = 500)
    from oracle import port
    port.load(native=True)
    X, y = synth(KIN, npts, 7); Z = inducing(KIN); ell = np.full(KIN["D"], KIN["ell"])
    for _ in range(args.warmup):
        port.sweep(X, y, Z, KIN["variance"], ell, 1.0e4)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.sweep(X, y, Z, KIN["variance"], ell, 1.0e4)
    dt = (time.perf_counter() - t0) / args.steps
    val = npts / dt
    line = {"impl": "reference", "metric": "vsgp_sweep_data_points_per_sec", "value": val, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "kin40k-shape VSGP sweep: D=8, M=512, SE-ARD, Float64; reference per-point schedule",
                       "sample": "%d points per step (one reference mini-batch)" % npts},
            "cpu_baseline": {"value": val, "unit": "points/s", "cores": 1, "kind": "port",
                             "sample": "%d-point mini-batch per step; C port of GPnode/UniSGPnode.jl:144-158 + :62-63 (Julia not installed; "
                                       "the per-point message-passing schedule is sequential: 1 thread)" % npts},
            "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-synthetic", action="store_true", help="skip the N=10M / M=1024 leg")
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

    ctx = SGPContext(local)
    if world > 1:
        uid = [SGPContext.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])


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
        ms_step_loc, ms_main_loc = ctx.sweep_timed_flushed(args.steps, 256)
        barrier()
        t_wall = time.perf_counter() - t_wall0
    ms_step = max_over_ranks(ms_step_loc)
    ms_main = max_over_ranks(ms_main_loc)
    info = ctx.last_sweep_info()
    value = world * cfg["N"] / (ms_step * 1e-3)

    # e2e through the C ABI with HOST buffers in pinned memory (sgp_pinned_alloc): per step the H2D copy of X / y
    # and the D2H read of Psi0 / Psi1 / Psi2 / sum_y2 (one sgp_sweep_psi_host call) are inside the timed region
    from gaussianprocessnode_b200 import pinned_empty
    Xp = pinned_empty(X.shape); Xp[...] = X
    yp = pinned_empty(y.shape); yp[...] = y
    psi1p = pinned_empty((cfg["M"],)); psi2p = pinned_empty((cfg["M"], cfg["M"]), order="F")
    for _ in range(3):
        ctx.sweep_psi_host(Xp, yp, out=(psi1p, psi2p))
    barrier()
    n_e2e = max(10, min(args.steps, 50))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        out = ctx.sweep_psi_host(Xp, yp, out=(psi1p, psi2p))     # sgp_sweep_psi_host: H2D of X / y, sweep, D2H of the statistics
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / n_e2e)
    h2d = X.nbytes + y.nbytes
    d2h = out[2].nbytes + out[1].nbytes + 32
    e2e_val = world * cfg["N"] / e2e_s

    peak, peak_src = fp64_peak()
    flops = cfg["N"] * cfg["M"] * (cfg["M"] + 1)
    achieved = flops / (ms_main * 1e-3) * 1e-12
    line = {
        "metric": "vsgp_sweep_data_points_per_sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "kin40k-shape VSGP sweep (BASELINE.json configs[1]): N=10000 points per GPU, D=8, M=512, SE-ARD, Float64; "
                               "Psi0/Psi1/Psi2 per step" + ("; statistics summed over the ranks inside the sweep kernel (NVLink peer memory, two-shot)" if world > 1 else ""),
                   "l2": "256 MB buffer rewritten on the stream before every timed step (inputs are smaller than L2); flush outside the timed intervals",
                   "parallelism": "N sharded over %d GPU(s)" % world, "wall_s_timed_region": t_wall},
        "e2e": {"value": e2e_val, "unit": "points/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(info["launches"] * args.steps),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic("kin40k"),
                     "kernel": "sweep4_kernel<128,32,8,256> (generate-once sweep: one cooperative launch per sweep) grid=%d block=%d smem=%d" % (info["grid"], info["block"], info["smem_bytes"]),
                     "algorithmic_flops_per_launch": flops, "ms_per_launch": ms_main, "peak_source": peak_src,
                     "note": "kin40k shape is 2.6 GFLOP: launch/latency-bound (66 us at peak); see synthetic_10M for the throughput-bound case"},
        "clocks": clk.summary(),
    }

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
        line["synthetic_10M"] = {
            "workload": "BASELINE.json configs[4]: N=10M (sharded: %d per GPU), D=8, M=1024, SE-ARD, Float64" % n_pad, "scaling": "strong",
            "value": world * n_pad / (ms_s * 1e-3), "unit": "points/s", "ms_per_step": ms_s, "steps": args.syn_steps,
            "psi2_tflops_all_gpus": world * fl / (ms_s * 1e-3) * 1e-12,
            "roofline": {"bound": "tensor", "achieved": fl / (ms_m * 1e-3) * 1e-12, "peak": peak, "unit": "TFLOP/s",
                         "frac": fl / (ms_m * 1e-3) * 1e-12 / peak, "traffic": traffic("synthetic"), "ms_per_launch": ms_m,
                         "algorithmic_flops_per_launch": fl, "kernel": "sweep4_kernel<128,32,8,256> grid=%d" % info2["grid"],
                         "peak_source": peak_src},
            "clocks": clk2.summary()}
        # the same sweep with 32 MB panels (ring of 96 MB: fewer slab turnovers, but the K_uf panels no longer stay in L2 -- they are
        # written back to and partly re-read from HBM); reported beside the default, not instead of it
        try:
            os.environ["SGP_SWEEP_SLAB_MB"] = "32"
            for _ in range(2):
                ctx.sweep_timed(1)
            barrier()
            ts2 = [ctx.sweep_timed(1)[1] for _ in range(args.syn_steps)]
            ms2 = max_over_ranks(float(np.mean(ts2)))
            line["synthetic_10M"]["slab_32MB_spilling"] = {"ms_per_launch": ms2, "tflops": fl / (ms2 * 1e-3) * 1e-12,
                                                           "frac": fl / (ms2 * 1e-3) * 1e-12 / peak}
        finally:
            os.environ.pop("SGP_SWEEP_SLAB_MB", None)
        del Xd, yd

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
            for _ in range(3):
                ctx.theta_objective(vb, Uvb, 1.0e4, 1e-8)
            t0 = time.perf_counter()
            for _ in range(10):
                ctx.theta_objective(vb, Uvb, 1.0e4, 1e-8)
            dt = (time.perf_counter() - t0) / 10
            from oracle import theta as otheta
            t0 = time.perf_counter(); otheta.neg_log_backwardmess_fast(cfg["variance"], ellb, yb, Xb, vb, Uvb, 1.0e4, Zb, 0, 1e-8)
            dt_cpu = time.perf_counter() - t0
            line["theta_step"] = {"workload": "objective + exact gradient of the theta step, one kin40k mini-batch: N=500, D=8, M=512 (host buffers in, "
                                              "D+2 scalars out; sgp_theta_objective)", "ms_per_call": dt * 1e3,
                                  "cpu_value_only_ms": dt_cpu * 1e3,
                                  "cpu_kind": "oracle restatement of neg_log_backwardmess_fast (value only, 1 thread; the reference adds ForwardDiff: "
                                              "3 chunked dual-number passes for 9 parameters)"}
            # one full mini-batch step of the kin40k driver (regression_kin40k.ipynb:196-228), device-resident streaming:
            # upload 500 points -> sweep -> posterior on the resident prior (becomes the next prior) -> theta objective + gradient
            ctx.kuu_factor(0.0, fetch=False)
            ctx.prior_set_isotropic(50.0)
            def mb_step():
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
            line["minibatch_step"] = {"workload": "kin40k driver mini-batch (500 points, M=512): H2D of the batch, sweep, posterior on the resident "
                                                  "streaming prior, theta objective + gradient; only D+2 scalars return to the host",
                                      "ms_per_minibatch": dt_mb * 1e3, "points_per_s": 500 / dt_mb,
                                      "reference": "~400 points/s end to end (3h30 for 500 epochs x 10000 points, regression_kin40k.ipynb:239; "
                                                   "author's Mac, includes RxInfer scheduling)"}
        except Exception as e:  # pragma: no cover
            line["theta_step"] = {"error": str(e)}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    if rank == 0 and world == 1:
        rate, secs = cpu_port_rate(KIN, KIN["N"], native=True, reps=5)
        line["cpu_baseline"] = {"value": rate, "unit": "points/s", "cores": 1, "kind": "port",
                                "sample": "all 10000 kin40k-shape points, best of 5 passes (%.1f s each); C port of the reference's "
                                          "per-point rule + prod schedule (oracle/sweep_port.c), Julia not installed" % secs}
        try:
            from oracle import batched
            Xs, ys = synth(KIN, KIN["N"], 7)
            t0 = time.perf_counter(); batched.psi_stats_point(Xs, ys, Z, KIN["variance"], ell); dt = time.perf_counter() - t0
            line["cpu_best_effort"] = {"value": KIN["N"] / dt, "unit": "points/s", "cores": os.cpu_count(),
                                       "kind": "numpy batched K_uf + OpenBLAS dgemm (not the reference's schedule)"}
        except Exception as e:  # pragma: no cover
            line["cpu_best_effort"] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
