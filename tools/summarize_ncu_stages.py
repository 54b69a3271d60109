"""Per-kernel summary of an `ncu --set full` capture of the HBM-bound stages (tools/prof_path.py) -> profiles/<tag>_ncu_stages.csv
usage: python tools/summarize_ncu_stages.py <report.ncu-rep> <tag> [hbm_gbs]
One row per kernel (launches averaged): duration, DRAM bytes read / written, achieved DRAM GB/s and its fraction of the
measured HBM copy bandwidth (MEASURED_PEAKS.json), L2 hit rate, registers, achieved occupancy, top stall."""
import collections
import csv
import json
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    hbm = float(sys.argv[3]) if len(sys.argv) > 3 else json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6650.0
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9,
         "second": 1.0, "s": 1.0}


def val(row, name):
    i = ix.get(name)
    if i is None or row[i] == "":
        return float("nan")
    return float(row[i].replace(",", "")) * scale.get(units[i], 1.0)


agg = collections.OrderedDict()
for row in rr[2:]:
    name = row[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("mcb::", "")
    a = agg.setdefault(name, collections.defaultdict(float))
    a["n"] += 1
    for k, m in (("dur", "gpu__time_duration.sum"), ("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum"),
                 ("l2hit", "lts__t_sector_hit_rate.pct"), ("regs", "launch__registers_per_thread"),
                 ("occ", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                 ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("issue", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
                 ("lsb", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct")):
        a[k] += val(row, m)
os.makedirs(os.path.join(root, "profiles"), exist_ok=True)
out = os.path.join(root, "profiles", f"{tag}_ncu_stages.csv")
with open(out, "w") as fh:
    fh.write("kernel,launches,ms_per_launch,dram_read_GB,dram_write_GB,achieved_dram_GBps,frac_of_measured_hbm,ncu_dram_pct_of_peak,l2_hit_pct,regs,achieved_occupancy_pct,issue_active_pct,long_scoreboard_stall_pct\n")
    for name, a in agg.items():
        n = a["n"]
        dur, rd, wr = a["dur"] / n, a["rd"] / n, a["wr"] / n
        gbps = (rd + wr) / dur / 1e9
        line = f'"{name}",{int(n)},{dur * 1e3:.3f},{rd / 1e9:.3f},{wr / 1e9:.3f},{gbps:.0f},{gbps / hbm:.3f},{a["dram_pct"] / n:.1f},{a["l2hit"] / n:.1f},{a["regs"] / n:.0f},{a["occ"] / n:.1f},{a["issue"] / n:.1f},{a["lsb"] / n:.1f}'
        fh.write(line + "\n")
        print(line)
print("->", out, f"(HBM peak used: {hbm} GB/s)")
