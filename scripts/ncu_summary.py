#!/usr/bin/env python
"""One line per captured kernel from an .ncu-rep (needs `ncu` on PATH):
duration, DRAM bytes, DRAM throughput %, achieved occupancy, registers, top stall reasons."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; units = rows[1]
def col(name):
    return hdr.index(name) if name in hdr else None
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    if not r: continue
    d = {}
    for w in want:
        c = col(w)
        if c is not None:
            d[w] = (r[c], units[c])
    name = d.get("Kernel Name", ("?",))[0].split("(")[0]
    def f(k):
        v = d.get(k)
        return float(v[0].replace(",", "")) if v and v[0] not in ("", "n/a") else float("nan")
    def bytes_(k):
        v = d.get(k)
        if not v: return float("nan")
        x = float(v[0].replace(",", "")); u = v[1]
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    dur = f("gpu__time_duration.sum"); du = d["gpu__time_duration.sum"][1]
    dur_us = dur / 1e3 if du.startswith("n") else dur * 1e3 if du.startswith("m") else dur
    rd, wr = bytes_("dram__bytes_read.sum"), bytes_("dram__bytes_write.sum")
    st = sorted(((float(r[hdr.index(h)].replace(",", "") or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:4]
    print("%-40s %8.1f us  dram %7.1f MB rd %7.1f MB wr (%.0f GB/s, %s%% of peak)  occ %s%%  regs %s  grid %s x %s  L1hit %s%% L2hit %s%%  stalls: %s" % (
        name[:40], dur_us, rd / 1e6, wr / 1e6, (rd + wr) / dur_us / 1e3, d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ("?",))[0],
        d.get("sm__warps_active.avg.pct_of_peak_sustained_active", ("?",))[0], d.get("launch__registers_per_thread", ("?",))[0],
        d.get("launch__grid_size", ("?",))[0], d.get("launch__block_size", ("?",))[0],
        d.get("l1tex__t_sector_hit_rate.pct", ("?",))[0], d.get("lts__t_sector_hit_rate.pct", ("?",))[0],
        ", ".join("%s %.1f" % (n, v) for v, n in st)))
