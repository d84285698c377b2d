"""profiles/r2_gemm_traffic.json from the `ncu --set full` capture of scripts/gemm_i8_bench.py --one (run on the GPU box
right after the capture, so the SHA-256 is that of the csrc/gemm.cu the captured library was built from).

    python scripts/make_gemm_traffic.py gpurun_out/r2_gemm_final_dram.csv gpurun_out/r2_gemm_traffic.json
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(src, dst, gemm_src=None):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    to_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    launches = []
    for r in data:
        rd = float(r[col["dram__bytes_read.sum"]]) * to_b[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * to_b[units[col["dram__bytes_write.sum"]]]
        launches.append({"kernel": r[col["Kernel Name"]].split("(")[0], "dram_read": int(rd), "dram_write": int(wr),
                         "us": float(r[col["gpu__time_duration.sum"]])})
    # launch order of gemm_i8_bench.py --one: (i8, f16) x (fp16 out, fp32 + residual) on 16384 x 1280 x 1280
    m, n, k = 16384, 1280, 1280
    f16 = launches[3]
    alg = m * k * 2 + n * k * 2 + 2 * m * n * 4
    gemm_src = gemm_src or os.path.join(ROOT, "dgq_b200", "csrc", "gemm.cu")
    out = {"gemm_cu_sha16": hashlib.sha256(open(gemm_src, "rb").read()).hexdigest()[:16],
           "dram_bytes": f16["dram_read"] + f16["dram_write"],
           "note": f"ncu --set full --clock-control none, one kind::f16 launch of the dominant plain fp32 + residual class "
                   f"(16384 x 1280 x 1280, cold L2): DRAM read {f16['dram_read'] / 1e6:.1f} MB + write {f16['dram_write'] / 1e6:.1f} MB "
                   f"vs {alg / 1e6:.1f} MB algorithmic (fp16 A and B once, fp32 residual in, fp32 result out; part of the "
                   f"result is still dirty in L2 when the kernel ends)",
           "algorithmic_bytes": alg, "launches": launches}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("gemm_cu_sha16", "dram_bytes", "algorithmic_bytes")}))


if __name__ == "__main__":
    main(*sys.argv[1:])
