"""Summarise an .ncu-rep: headline metrics + instruction mix (needs ncu on PATH). Usage: ncu_summary.py file.ncu-rep"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print(f"  {h:90s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = 0; byop = collections.Counter(); samp = collections.Counter()
for r in rows[hi + 1:]:
    try: n = int(r[iE])
    except Exception: continue
    t = r[iS].split()
    if not t: continue
    op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
    byop[op] += n; tot += n; samp[op] += int(r[iSm] or 0)
print("total warp instructions", tot)
stot = sum(samp.values()) or 1
for op, n in byop.most_common(22):
    print(f"  {op:10s} {n:13d} {100*n/tot:5.1f}%   stall-samples {100*samp[op]/stot:5.1f}%")
