"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep itself stays in gpurun_out/).

  python scripts/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/rNN_kernel_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    acc = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}[row["Metric Unit"]]
        base = re.sub(r"\(.*", "", row["Kernel Name"])
        acc[base][0] += 1
        acc[base][1] += v
    tot = sum(v[1] for v in acc.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: one step, {sum(v[0] for v in acc.values())} launches, "
          f"{tot:.3f} ms summed (cold-cache, serialised: compare SHARES)")
    print(f"{'ms':>10s} {'share':>6s} {'calls':>6s}  kernel")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.3f} {100 * v[1] / tot:5.1f}% {v[0]:6d}  {k[:120]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none, {len(data)} captured launches of {path}")
    for r in data:
        print("kernel:", r[idx["Kernel Name"]][:100], "grid", r[idx["Grid Size"]], "block", r[idx["Block Size"]])
        for k in KEYS:
            if k in idx:
                print(f"  {k:78s} {r[idx[k]]:>16s} {units[idx[k]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
