"""Summarise an ncu report (read here, no GPU): key pipe/memory metrics, stall reasons, per-phase sample split.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "?")[:120])
    for k in KEYS:
        if k in d:
            print("  %-95s %-10s %s" % (k, units[hdr.index(k)], d[k]))
    stalls = sorted(((float(v), k) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued") and v),
                    reverse=True)
    tot = sum(s for s, _ in stalls) or 1.0
    print("  stall reasons (share of samples):")
    for s, k in stalls[:8]:
        print("    %-40s %5.1f %%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * s / tot))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
data = []
for r in rows:
    if r and r[0] == "Address":
        h = r
        continue
    if h and len(r) == len(h):
        try:
            data.append((r[h.index("Source")], int(r[h.index("Warp Stall Sampling (All Samples)")]), int(r[h.index("Instructions Executed")]),
                         int(r[h.index("L1 Wavefronts Shared")] or 0), int(r[h.index("L1 Wavefronts Shared Ideal")] or 0)))
        except ValueError:
            pass
if data:
    tot = sum(d[1] for d in data) or 1
    print("per-phase split of warp-stall samples (phases delimited by BAR.SYNC):")
    start = 0
    bars = [i for i, d in enumerate(data) if "BAR.SYNC" in d[0]] + [len(data) - 1]
    for i, b in enumerate(bars):
        seg = data[start:b + 1]
        s = sum(d[1] for d in seg)
        kinds = {}
        for d in seg:
            op = d[0].split()[0] if not d[0].startswith("@") else d[0].split()[1]
            kinds[op] = kinds.get(op, 0) + d[2]
        top = ", ".join("%s %d" % (k, v) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])[:5])
        wf = sum(d[3] for d in seg); wfi = sum(d[4] for d in seg)
        print("  phase %d: %5.1f %% of samples, %d SASS instr, smem wavefronts %d (ideal %d); executed: %s" % (i, 100 * s / tot, len(seg), wf, wfi, top))
        start = b + 1
    print("top 12 instructions by samples:")
    for d in sorted(data, key=lambda d: -d[1])[:12]:
        print("  %5.1f %%  %s" % (100 * d[1] / tot, d[0][:90]))
