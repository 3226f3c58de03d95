#!/usr/bin/env python
"""Turn the raw artefacts a gpurun call brought back (gpurun_out/) into the small text summaries that are
committed under profiles/ (per round).  Usage: python tools/summarize_profiles.py r01"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)


def launches(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        agg[row["Kernel Name"][:90]].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none): cold-cache, serialised -> compare SHARES\n")
        f.write("# source: %s\n" % path)
        f.write("%-92s %5s %11s %9s %9s %9s %7s\n" % ("kernel", "n", "sum_us", "mean_us", "min_us", "max_us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-92s %5d %11.1f %9.1f %9.1f %9.1f %6.1f%%\n" % (k, len(v), sum(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def rep(path, out, note=""):
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s\n# %s\n" % (path, note))
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            f.write("\n## %s   (launch id %s)\n" % (d.get("Kernel Name"), d.get("ID")))
            for k in KEYS:
                if k in d:
                    f.write("%-100s %s %s\n" % (k, d[k], u[k]))


for name in os.listdir(G):
    if name.startswith("launches") and name.endswith(".csv"):
        launches(os.path.join(G, name), os.path.join(P, "%s_%s.txt" % (tag, name[:-4])))
    if name.endswith(".ncu-rep"):
        rep(os.path.join(G, name), os.path.join(P, "%s_%s.txt" % (tag, name[:-8])))
    if name.startswith("bench") and name.endswith(".json"):
        try:
            d = json.load(open(os.path.join(G, name)))
            json.dump(d, open(os.path.join(P, "%s_%s" % (tag, name)), "w"), indent=1)
        except Exception:
            pass
print("profiles written:", sorted(os.listdir(P)))
