"""Digest of an ncu --set full report: headline metrics per launch, opcode mix and the most-sampled
instructions with their stall reasons (reads the .ncu-rep here, no GPU needed).

  python scripts/ncu_digest.py gpurun_out/x.ncu-rep [launch index ...]
"""
import collections
import csv
import io
import subprocess
import sys

HEAD = ["gpu__time_duration.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
STALLS = ["stall_long_sb", "stall_wait", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_math",
          "stall_short_sb", "stall_barrier", "stall_mio", "stall_no_inst", "stall_lg", "stall_dispatch", "stall_sleep",
          "stall_membar", "stall_tex"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    which = [int(a) for a in sys.argv[2:]]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for li, r in enumerate(data):
        if which and li not in which:
            continue
        print(f"== launch {li}: {r[idx['Kernel Name']][:90]} grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for k in HEAD:
            if k in idx:
                print(f"   {k:82s} {r[idx[k]]:>14s} {units[idx[k]]}")
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--launch-skip", str(li),
                                                   "--launch-count", "1"]))))
        h2 = next(i for i, x in enumerate(src) if x and x[0] == "Address")
        sh = src[h2]
        sd = [x for x in src[h2 + 1:] if len(x) == len(sh) and x[sh.index('# Samples')].isdigit()]
        ia, ie, isamp = sh.index("Source"), sh.index("Instructions Executed"), sh.index("# Samples")
        tot = sum(int(x[ie]) for x in sd) or 1
        tots = sum(int(x[isamp]) for x in sd) or 1
        ops, ops_s = collections.Counter(), collections.Counter()
        for x in sd:
            t = [o for o in x[ia].split() if not o.startswith("@")]
            op = t[0].split(".")[0] if t else ""
            ops[op] += int(x[ie])
            ops_s[op] += int(x[isamp])
        print(f"   executed warp-instructions {tot}, samples {tots}")
        print("   opcode mix: " + ", ".join(f"{o} {100 * c / tot:.1f}% ({100 * ops_s[o] / tots:.0f}% smp)" for o, c in ops.most_common(12)))
        st = {c: sum(int(x[sh.index(c)]) for x in sd) for c in STALLS if c in sh}
        print("   stall samples: " + ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v))
        top = sorted(range(len(sd)), key=lambda i: -int(sd[i][isamp]))[:14]
        for i in sorted(top):
            x = sd[i]
            why = max(((int(x[sh.index(c)]), c[6:]) for c in STALLS if c in sh), default=(0, ""))
            print(f"     [{i:5d}] smp {x[isamp]:>5s} exec {x[ie]:>9s} {why[1]:>16s}  {x[ia].strip()[:80]}")


if __name__ == "__main__":
    main()
