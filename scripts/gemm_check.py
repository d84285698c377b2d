"""Correctness + timing of dgq_gemm_f16 on the dominant UNet GEMM shapes.  DGQ_GEMM_CTAS=1|2 pins the
kernel variant (read once per process).  CUDA events on the launch stream, L2 flushed between runs."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgq_b200 import ops

CHECK = [(128, 128, 64), (256, 320, 128), (100, 8, 72), (154, 640, 768), (16, 1280, 320), (1232, 640, 2048),
         (4096, 320, 2880), (1024, 1280, 1280), (300, 2560, 640), (8192, 1280, 640), (16384, 10240, 1280)]
BENCH = [(16384, 10240, 1280), (16384, 1280, 1280), (16384, 1280, 5120), (16384, 1280, 11520),
         (65536, 5120, 640), (65536, 640, 640), (65536, 640, 5760), (65536, 640, 2560), (262144, 320, 2880),
         (4096, 320, 2880), (16384, 320, 2880)]


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


def main():
    dev = "cuda"
    tag = os.environ.get("DGQ_GEMM_CTAS", "auto")
    ok = True
    for m, n, k in CHECK:
        g = torch.Generator(device=dev).manual_seed(m + n + k)
        a = (torch.randn(m, k, generator=g, device=dev) * 0.7).half()
        b = torch.randint(-15, 16, (n, k), generator=g, device=dev).half()
        scale = torch.rand(n, generator=g, device=dev) * 0.01 + 0.001
        bias = torch.randn(n, generator=g, device=dev) * 0.1
        rows = 64 if m % 64 == 0 else m
        temb = torch.randn(m // rows, n, generator=g, device=dev)
        resid = torch.randn(m, n, generator=g, device=dev)
        rs = torch.rand(rows, generator=g, device=dev) + 0.5
        ref = (a.float() @ b.float().t()) * rs.repeat(m // rows)[:, None] * scale + bias \
            + temb.repeat_interleave(rows, 0) + resid
        out = ops.gemm(a, b, n, scale=scale, bias=bias, temb=temb, rows_per_batch=rows, resid=resid, want_f32=True,
                       row_scale=rs, row_period=rows)
        out2 = ops.gemm(a, b, n, scale=scale, bias=bias, temb=temb, rows_per_batch=rows, resid=resid, want_f32=True,
                        row_scale=rs, row_period=rows)
        e = rel(out, ref)
        same = torch.equal(out, out2)
        good = e < 1e-5 and same
        ok &= good
        print(f"[{tag}] check {m}x{n}x{k}: rel {e:.2e} repeatable {same} {'ok' if good else 'FAIL'}", flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = []
    for m, n, k in BENCH:
        a = torch.randn(m, k, device=dev).half()
        b = torch.randint(-15, 16, (n, k), device=dev).half()
        scale = torch.rand(n, device=dev)
        bias = torch.rand(n, device=dev)
        resid = torch.randn(m, n, device=dev)
        out = torch.empty(m, n, dtype=torch.float32, device=dev)
        for variant, kw in (("f32out+bias", dict(scale=scale, bias=bias)), ("f32out+bias+resid", dict(scale=scale, bias=bias, resid=resid))):
            for _ in range(2):
                ops.gemm(a, b, n, out=out, **kw)
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                ops.gemm(a, b, n, out=out, **kw)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            res.append(dict(ctas=tag, m=m, n=n, k=k, epilogue=variant, ms=round(t, 4), tflops=round(2.0 * m * n * k / t / 1e9, 1)))
            print(res[-1], flush=True)
        del a, b, resid, out
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/gemm_check_{tag}.json", "w"), indent=1)
    print("ALL OK" if ok else "FAILURES")


if __name__ == "__main__":
    main()
