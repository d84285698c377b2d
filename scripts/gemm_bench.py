"""Time dgq_gemm_f16 on the dominant SDXL/SD GEMM shapes (CUDA events, L2 flushed between runs)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgq_b200 import ops

SHAPES = [(16384, 10240, 1280), (16384, 1280, 1280), (16384, 1280, 5120), (16384, 1280, 11520),
          (65536, 5120, 640), (65536, 640, 640), (65536, 640, 5760), (4096, 320, 2880), (16384, 320, 2880)]


def main():
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = []
    for m, n, k in SHAPES:
        a = torch.randn(m, k, device=dev).half()
        b = torch.randint(-15, 16, (n, k), device=dev).half()
        scale = torch.rand(n, device=dev)
        out = torch.empty(m, n, dtype=torch.float16, device=dev)
        for _ in range(3):
            ops.gemm(a, b, n, scale=scale, out=out)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            ops.gemm(a, b, n, scale=scale, out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        tf = 2.0 * m * n * k / t / 1e9
        # cuBLAS (library) reference point for the same shape
        bt = b.t().contiguous()
        for _ in range(3):
            torch.matmul(a, bt)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        flush.zero_()
        e0.record()
        torch.matmul(a, bt)
        e1.record()
        torch.cuda.synchronize()
        tl = e0.elapsed_time(e1)
        res.append(dict(m=m, n=n, k=k, ms=t, tflops=tf, cublas_ms=tl, cublas_tflops=2.0 * m * n * k / tl / 1e9))
        print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
