"""Summarise ncu outputs into small text files for profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv  > profiles/<name>.txt
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/<name>.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("vcsmc::<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none : per-kernel totals over the captured launches")
    print("# (cold-cache, serialised: compare SHARES, not absolutes)   source: %s" % path)
    print("%-72s %8s %12s %7s" % ("kernel", "launches", "total ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %8d %12.3f %6.1f%%" % (k[-72:], v[0], v[1], 100 * v[1] / tot))
    print("%-72s %8d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none --import-source on   source: %s" % path)
    for r in rows[2:]:
        print("kernel: %s" % r[idx["Kernel Name"]])
        for k in KEYS:
            if k in idx:
                print("  %-72s %s %s" % (k, r[idx[k]], units[idx[k]]))
        st = [(h, float(r[i])) for h, i in idx.items()
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and r[i]]
        for h, v in sorted(st, key=lambda x: -x[1])[:7]:
            print("  stall %-66s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
