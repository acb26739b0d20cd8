"""Summarise an .ncu-rep (first kernel): key metrics, instruction mix, hot SASS segments, top stall sites.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    if k in M: print(f"{k:95s} {M[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(hdr)}
byop = collections.Counter(); tot = 0; items = []
wf = collections.defaultdict(lambda: [0, 0])
for k, r in enumerate(data):
    try: n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    except Exception: continue
    t = r[ix['Source']].strip().split(); op = (t[1] if t[0].startswith('@') else t[0])
    byop[op.split('.')[0]] += n; tot += n
    items.append((s, k, n, r[ix['Source']].strip()[:80], r[ix['stall_long_sb']], r[ix['stall_short_sb']], r[ix['stall_wait']], r[ix['stall_math']], r[ix['stall_mio']]))
    try:
        w = int(r[ix['L1 Wavefronts Shared']]); wi = int(r[ix['L1 Wavefronts Shared Ideal']])
        if w: wf[op][0] += w; wf[op][1] += wi
    except Exception: pass
print("total SASS executed", tot)
print("mix:", ", ".join(f"{o} {100*n/tot:.1f}%" for o, n in byop.most_common(18)))
print("smem wavefronts (actual/ideal):", {k: tuple(v) for k, v in wf.items()})
items.sort(reverse=True)
print("top stall sites: samples idx exec source long_sb short_sb wait math mio")
for it in items[:22]: print(it)
