"""Turn gpurun_out/{launches_TAG.csv,kernels_TAG.ncu-rep} into the tracked text summaries under profiles/.

    python scripts/summarize_profiles.py r1
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
]


def launches():
    path = os.path.join(ROOT, "gpurun_out", f"launches_{TAG}.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(OUT, f"launches_{TAG}.txt"), "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (bench.py --steps 2 --warmup 3 --samples 1024 --quick --no-e2e --no-cpu-baseline; all launches incl. warm-up and the per-kernel timing legs)\n")
        fh.write(f"# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  total {total/1e6:.3f} ms\n")
        fh.write(f"{'kernel':80s} {'launches':>8s} {'mean_us':>10s} {'total_ms':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write(f"{k[:80]:80s} {len(v):8d} {sum(v)/len(v)/1e3:10.2f} {sum(v)/1e6:10.3f} {100*sum(v)/total:6.1f}%\n")


def kernels():
    rep = os.path.join(ROOT, "gpurun_out", f"kernels_{TAG}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    seen = set()
    traffic = {}
    with open(os.path.join(OUT, f"kernels_{TAG}.txt"), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on (one capture per kernel; first instance of each shown)\n")
        for n, r in enumerate(rows[2:]):
            name = r[idx["Kernel Name"]]
            if name in seen:
                continue
            seen.add(name)
            try:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
                wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
                short = name.replace("void ", "").split("<")[0].split("(")[0].split("::")[-1]
                traffic[short] = rd + wr
            except Exception:
                pass
            fh.write(f"\n== {name}\n")
            for m in METRICS:
                if m in idx:
                    fh.write(f"   {m:70s} {r[idx[m]]:>16s} {units[idx[m]]}\n")
            st = sorted(((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)
            fh.write("   top stalls (warps per issue-active cycle): " + ", ".join(f"{nm}={v:.2f}" for v, nm in st[:6]) + "\n")
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(n), "--launch-count", "1"],
                                 capture_output=True, text=True).stdout
            srows = list(csv.reader(src.splitlines()))
            if len(srows) > 3:
                h2 = srows[2]
                try:
                    ii, si = h2.index("Instructions Executed"), h2.index("# Samples")
                except ValueError:
                    continue
                items = []
                for rr in srows[3:]:
                    if rr and rr[0].isdigit():
                        try:
                            items.append((int(rr[ii]), int(rr[si]), int(rr[0]), rr[1].strip()))
                        except ValueError:
                            pass
                ti = sum(i[0] for i in items) or 1
                ts = sum(i[1] for i in items) or 1
                fh.write("   hottest source lines (share of warp instructions / of stall samples):\n")
                for inst, samp, line, text in sorted(items, reverse=True)[:8]:
                    fh.write(f"     {100*inst/ti:5.1f}% {100*samp/ts:5.1f}%  L{line:<4d} {text[:100]}\n")
    import json
    with open(os.path.join(OUT, f"traffic_{TAG}.json"), "w") as fh:
        json.dump(traffic, fh, indent=1)


if __name__ == "__main__":
    launches()
    kernels()
    print(os.listdir(OUT))
