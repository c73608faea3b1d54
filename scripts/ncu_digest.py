"""Digest of one `ncu --set full` report: headline counters, top stalls, per-source-line instruction/stall shares."""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.008
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, r = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    if w in hdr:
        print(f"{w:70s} {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
print("stalls:", ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[2]
ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
out = []
for r in rows[3:]:
    if r[0] and r[0].isdigit():
        try:
            out.append((int(r[0]), int(r[ii]), int(r[isamp]), r[1].strip()))
        except ValueError:
            pass
tot = sum(o[1] for o in out) or 1
ts = sum(o[2] for o in out) or 1
print(rows[1][1][:100], "warp-instr", tot, "samples", ts)
for l, i, s, text in sorted(out):
    if i / tot > thr or s / ts > thr:
        print(f"L{l:<4} {100*i/tot:5.1f}% {100*s/ts:5.1f}%  {text[:110]}")
