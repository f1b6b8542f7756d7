"""Summarise ncu outputs into the text files committed under profiles/.
  python tools/ncu_summary.py launches <launches.csv>          -> per-kernel count / total / share
  python tools/ncu_summary.py raw <report.ncu-rep> [regex]      -> key counters of each captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("<unnamed>::", "")
        v = float(d["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(d["Metric Unit"], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print("%-48s %8s %12s %8s %10s" % ("kernel", "launches", "total ms", "share", "avg ms"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %8d %12.3f %7.1f%% %10.4f" % (k[:48], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    print("%-48s %8d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


def raw(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    units = rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  (id %s)" % (d.get("Kernel Name", "?")[:90], d.get("ID")))
        for k in hdr:
            if pattern and re.search(pattern, k) or (not pattern and k in KEYS):
                print("   %-75s %18s %s" % (k, d[k], units[hdr.index(k)]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
