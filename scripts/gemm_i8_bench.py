"""kind::i8 vs kind::f16 qGEMM on the dominant SDXL / SD shapes (CUDA events, L2 flushed between runs): the measured
side of north_star (b)'s "i8 or f16" decision.  Writes gpurun_out/gemm_i8_bench.json (summarised under
profiles/r2_gemm_i8_vs_f16.txt).

    python scripts/gemm_i8_bench.py            # all shapes
    python scripts/gemm_i8_bench.py --one      # one launch of each kind on 16384x1280x1280 (for ncu)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dgq_b200 import ops  # noqa: E402

SHAPES = [(16384, 10240, 1280), (16384, 1280, 1280), (16384, 1280, 5120), (16384, 1280, 11520),
          (65536, 5120, 640), (65536, 640, 640), (65536, 640, 5760), (4096, 320, 2880), (1024, 1280, 1280),
          (256, 1280, 1280)]


def operands(m, n, k, wbits, dev):
    g = torch.Generator().manual_seed(m + n + k)
    a8 = torch.randint(0, 256, (m, k), generator=g, dtype=torch.int32).to(torch.uint8).to(dev)
    wl = 2 ** wbits
    codes = torch.randint(0, wl, (n, k), generator=g, dtype=torch.int32).to(torch.uint8).to(dev)
    wz = torch.full((n,), float(wl // 2), device=dev)
    b8, colsum, b_off = ops.weight_to_i8(codes, wz, n, float(wl - 1))
    az = torch.tensor([128.0], device=dev)
    ad = torch.tensor([0.02], device=dev)
    a16 = (a8.float() - 128.0).half()                      # the exact-integer kind::f16 operands of the same problem
    b16 = (codes.float() - float(wl // 2)).half()
    scale = (torch.rand(n, generator=g) * 0.01 + 0.001).to(dev)
    return a8, b8, colsum, b_off, az, ad, a16, b16, scale


def timed(fn, flush, reps=5):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = "cuda"
    one = "--one" in sys.argv
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = []
    for m, n, k in ([(16384, 1280, 1280)] if one else SHAPES):
        for wbits in ((8,) if one else (8, 4)):
            a8, b8, colsum, b_off, az, ad, a16, b16, scale = operands(m, n, k, wbits, dev)
            row = dict(m=m, n=n, k=k, wbits=wbits)
            for out_f32 in (False, True):
                resid = torch.randn(m, n, device=dev) if out_f32 else None
                out = torch.empty(m, n, dtype=torch.float32 if out_f32 else torch.float16, device=dev)
                f_i8 = lambda: ops.gemm(a8, b8, n, scale=scale, out=out, resid=resid, row_scale=ad, row_zp=az,  # noqa: E731
                                        colsum=colsum, b_off=b_off)
                f_16 = lambda: ops.gemm(a16, b16, n, scale=scale, out=out, resid=resid, row_scale=ad)  # noqa: E731
                if one:
                    f_i8(); f_16(); torch.cuda.synchronize()
                    continue
                t8, t16 = timed(f_i8, flush), timed(f_16, flush)
                y8 = out.clone(); f_16(); torch.cuda.synchronize()
                tag = "f32+resid" if out_f32 else "f16out"
                row[tag] = dict(i8_ms=round(t8, 4), f16_ms=round(t16, 4), i8_tops=round(2.0 * m * n * k / t8 / 1e9, 1),
                                f16_tflops=round(2.0 * m * n * k / t16 / 1e9, 1), speedup=round(t16 / t8, 3),
                                max_abs_diff=float((y8.float() - out.float()).abs().max()))
            if not one:
                res.append(row)
                print(json.dumps(row), flush=True)
    if not one:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open("gpurun_out/gemm_i8_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
