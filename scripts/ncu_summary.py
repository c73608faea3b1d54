"""Print the headline metrics, top stalls and hottest source lines of every distinct kernel in an .ncu-rep.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [n_lines]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 16
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
     "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
     "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
seen = set()
for n, r in enumerate(rows[2:]):
    name = r[idx["Kernel Name"]]
    if name in seen:
        continue
    seen.add(name)
    print("==", name[:70])
    for m in M:
        if m in idx:
            print(f"   {m:62s} {r[idx[m]]:>16s} {units[idx[m]]}")
    st = sorted(((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)
    print("   stalls:", ", ".join(f"{nm}={v:.2f}" for v, nm in st[:6]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(n), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    h2 = next((x for x in srows if "Instructions Executed" in x), None)
    if not h2:
        continue
    ii, si = h2.index("Instructions Executed"), h2.index("# Samples")
    items = []
    for rr in srows:
        if rr and rr[0].isdigit():
            try:
                items.append((int(rr[ii]), int(rr[si]), int(rr[0]), rr[1].strip()))
            except (ValueError, IndexError):
                pass
    ti, ts = sum(i[0] for i in items) or 1, sum(i[1] for i in items) or 1
    for inst, samp, line, text in sorted(items, reverse=True)[:nl]:
        print(f"     {100*inst/ti:5.1f}% {100*samp/ts:5.1f}%  L{line:<4d} {text[:105]}")
