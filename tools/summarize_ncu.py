"""Turn gpurun_out/<tag>_launches.csv (+ <tag>_prof_tc.ncu-rep) into the small summaries kept under profiles/.
usage: python tools/summarize_ncu.py <tag>"""
import csv
import collections
import os
import subprocess
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)

# ---- launch list: per-kernel count / total / share (cold-cache, serialised: compare shares) -----------
rows = []
with open(os.path.join(go, f"{tag}_launches.csv")) as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "s": 1e9, "second": 1e9}.get(unit, 1.0)
        name = r["Kernel Name"]
        short = name.split("(")[0].replace("void ", "").replace("mcb::", "")
        rows.append((short if ("mcb" in name or "tc_cor" in name or "_kernel" in short) else "torch/other: " + short[:60], ns))
agg = collections.OrderedDict()
for name, ns in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
with open(os.path.join(pr, f"{tag}_launch_list.csv"), "w") as fh:
    fh.write("kernel,launches,total_ms,share\n")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f'"{name}",{n},{ns / 1e6:.3f},{ns / tot:.4f}\n')
print("launch list:", len(rows), "launches,", f"{tot / 1e6:.1f} ms total")

# ---- full capture of the dominant kernel --------------------------------------------------------------
rep = os.path.join(go, f"{tag}_prof_tc.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rr[0], rr[1], rr[2]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
            "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
            "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum"]
    with open(os.path.join(pr, f"{tag}_ncu_tc_cor_kernel.csv"), "w") as fh:
        fh.write("metric,unit,value\n")
        for i, h in enumerate(hdr):
            if h in want or h.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"):
                fh.write(f'"{h}","{units[i]}","{vals[i]}"\n')
                print(h, units[i], vals[i])
