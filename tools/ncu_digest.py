"""Compact digest of one `ncu --set full` report: the numbers profiles/ncu_r2_summary.md quotes, and the top stall lines.
usage: python tools/ncu_digest.py gpurun_out/x.ncu-rep [n_lines]"""
import csv, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 6
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, u, v = r[0], r[1], r[2]
col = {n: i for i, n in enumerate(h)}
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.pct_of_peak_sustained_elapsed"]
for n in want:
    if n in col:
        print(f"{n:92s} {v[col[n]]} {u[col[n]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, x in enumerate(rows) if "# Samples" in x)
hdr = rows[hi]
si, ci, ei = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = [(i, x) for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
data = [x for x in rows[hi + 1:] if len(x) > max(si, ci, ei)]
tot = sum(float(x[si] or 0) for x in data) or 1.0
agg = {}
for x in data:
    for i, s in stall:
        agg[s] = agg.get(s, 0.0) + float(x[i] or 0)
ssum = sum(agg.values()) or 1.0
print("stall mix:", ", ".join(f"{k[6:]} {100 * a / ssum:.0f}%" for k, a in sorted(agg.items(), key=lambda t: -t[1])[:6]))
for val, x in sorted(((float(x[si] or 0), x) for x in data), key=lambda t: -t[0])[:top]:
    st = sorted(((float(x[i] or 0), s) for i, s in stall), reverse=True)[:2]
    print(f"  {100 * val / tot:5.1f}%  exec={x[ei]:>9}  {x[ci].strip()[:72]:72s} {[(s[6:], int(a)) for a, s in st]}")
