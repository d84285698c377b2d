"""Micro-benchmarks of the hot kernels at the headline workload's shapes (SDXL, batch 16).

  python scripts/kern_bench.py [--only ln,attn,prod,gemm] [--ncu]

CUDA events around back-to-back launches on inputs larger than L2; prints achieved GB/s (algorithmic
bytes) or TFLOP/s per case and writes gpurun_out/kern_bench.json.  --ncu: one launch per case, no
warm-up (for `ncu --set full -k regex:dgq`).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dgq_b200 import ops  # noqa: E402

DEV = "cuda"
NCU = "--ncu" in sys.argv
ONLY = None
for i, a in enumerate(sys.argv):
    if a == "--only":
        ONLY = set(sys.argv[i + 1].split(","))
RES = []


def timed(fn, reps=10):
    if NCU:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def kwise(c, g=16, bits=8, seed=0):
    gen = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, g, (c,), generator=gen)
    lo = -(torch.rand(g, generator=gen) * 3 + 1)
    hi = torch.rand(g, generator=gen) * 3 + 1
    d = (hi - lo) / (2 ** bits - 1)
    z = torch.round(-lo / d)
    return ops.qparam_from_ckpt(d[lab].view(1, 1, -1), z[lab].view(1, 1, -1), float(2 ** bits - 1), DEV)


def report(kind, label, ms, gbytes=None, tflop=None):
    r = dict(kind=kind, label=label, ms=round(ms, 4))
    if gbytes is not None:
        r["GBps"] = round(gbytes / (ms / 1e3), 1)
    if tflop is not None:
        r["TFLOPs"] = round(tflop / (ms / 1e3), 1)
    RES.append(r)
    print(r, flush=True)


def bench_ln():
    for m, c in ((16384, 1280), (65536, 640)):
        x = torch.randn(m, c, device=DEV)
        g, b = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.1
        for nq in (3, 1):
            qs = [kwise(c, seed=i) for i in range(nq)]
            ms = timed(lambda: ops.ln_quant(x, g, b, 1e-5, qs))
            report("ln", f"ln_quant m={m} c={c} nq={nq} kwise", ms, gbytes=m * c * (4 + 2 * nq) / 1e9)
    x = torch.randn(16 * 77, 2048, device=DEV)
    qs = [kwise(2048, seed=i) for i in range(2)]
    ms = timed(lambda: ops.row_quant(x, qs))
    report("ln", "row_quant ctx 1232x2048 nq=2", ms, gbytes=x.numel() * 8 / 1e9)


def bench_prod():
    for b, hw, c0, c1, up in ((16, 32, 1280, 0, False), (16, 64, 640, 0, False), (16, 128, 320, 0, False),
                             (16, 32, 1280, 1280, False), (16, 64, 1280, 0, True)):
        c = c0 + c1
        hs = hw // 2 if up else hw
        x0 = torch.randn(b, hs, hs, c0, device=DEV)
        x1 = torch.randn(b, hs, hs, c1, device=DEV) if c1 else None
        q = kwise(9 * c)
        mean, rstd = ops.gn_stats(x0, x1, b, hs * hs, 1e-5)
        gn = (mean, rstd, torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.1)
        if up:
            gn = None
        ms = timed(lambda: ops.act_producer(x0, src1=x1, batch=b, h=hw, w=hw, upsample=up, ksize=3, gn=gn,
                                            act=0 if up else 1, q=q, pad_quantized=True), reps=5)
        m = b * hw * hw
        report("prod", f"conv3x3 producer b={b} {hw}x{hw} c={c0}+{c1} up={up}", ms,
               gbytes=(b * hs * hs * c * 4 + m * 9 * c * 2) / 1e9)
        ms = timed(lambda: ops.gn_stats(x0, x1, b, hs * hs, 1e-5))
        report("prod", f"gn_stats b={b} {hs}x{hs} c={c}", ms, gbytes=b * hs * hs * c * 4 / 1e9)


def bench_attn():
    shapes = [(16, 10, 4096, 4096, 64, False, "sdxl self 64x64"), (16, 20, 1024, 1024, 64, False, "sdxl self 32x32"),
              (16, 10, 4096, 77, 64, True, "sdxl cross 64x64"), (16, 20, 1024, 77, 64, True, "sdxl cross 32x32")]
    if ONLY is not None and "attn1" in ONLY:
        shapes = shapes[1:2]
    for b, h, t, s, d, sp, label in shapes:
        dp = (d + 63) // 64 * 64
        x = torch.randn(b * t, h * d, device=DEV)
        kx = torch.randn(b * s, h * d, device=DEV)
        q = ops.qkv_pack(x, b, t, h, d, dp)
        k = ops.qkv_pack(kx, b, s, h, d, dp)
        v = ops.qkv_pack(kx, b, s, h, d, dp, transpose=True)
        qo = kwise(h * d)
        ms = timed(lambda: ops.attention(q, k, v, d, map_mode=ops.MAP_LOG2, real_time=True, start_peak=sp, out_q=qo),
                   reps=5)
        report("attn", label, ms, tflop=4.0 * b * h * t * s * d / 1e12)


def bench_gemm():
    shapes = [(16384, 10240, 1280, "geglu"), (16384, 1280, 1280, "resid"), (16384, 1280, 5120, "resid"),
              (16384, 1280, 11520, "resid"), (65536, 5120, 640, "geglu"), (65536, 640, 640, "resid"),
              (65536, 640, 2560, "resid"), (65536, 640, 5760, "resid"), (16384, 1280, 1280, "qkv"),
              (1232, 1280, 2048, "qkv"), (262144, 320, 2880, "resid")]
    if ONLY is not None and "gemm1" in ONLY:
        shapes = shapes[1:2] + shapes[5:6]
    for m, n, k, kind in shapes:
        a = torch.randn(m, k, device=DEV).half()
        w = torch.randint(-8, 8, (n, k), device=DEV).half()
        scale = torch.rand(n, device=DEV) * 0.01
        bias = torch.randn(n, device=DEV)
        if kind == "resid":
            resid = torch.randn(m, n, device=DEV)
            out = torch.empty(m, n, device=DEV)
            fn = lambda: ops.gemm(a, w, n, scale=scale, bias=bias, resid=resid, out=out)  # noqa: E731
        elif kind == "geglu":
            q2 = kwise(n // 2)
            fn = lambda: ops.gemm(a, w, n, scale=scale, bias=bias, epi=ops.EPI_GEGLU, q2=q2)  # noqa: E731
        else:
            heads, d = n // 64, 64
            tok = 1024 if m % 1024 == 0 else 77
            bb = m // tok
            q2 = kwise(d)
            dst = ops.qkv_dest(bb, tok, heads, d, 64, False, DEV)
            fn = lambda: ops.gemm(a, w, n, scale=scale, bias=bias, epi=ops.EPI_QKV, q2=q2, out=dst,  # noqa: E731
                                  qkv=(heads, d, 64, tok, (tok + 7) // 8 * 8, False, False))
        ms = timed(fn, reps=5)
        report("gemm", f"{kind} m={m} n={n} k={k}", ms, tflop=2.0 * m * n * k / 1e12)


def main():
    table = {"ln": bench_ln, "prod": bench_prod, "attn": bench_attn, "attn1": bench_attn, "gemm": bench_gemm, "gemm1": bench_gemm}
    for name, fn in table.items():
        if ONLY is None or name in ONLY:
            fn()
    os.makedirs("gpurun_out", exist_ok=True)
    if not NCU:
        json.dump(RES, open("gpurun_out/kern_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
