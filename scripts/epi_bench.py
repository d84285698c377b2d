"""Time the qGEMM epilogue modes on the SDXL shapes they run on (CUDA events, L2 flushed)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgq_b200 import ops

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def qp(n, mode):
    if mode == "none":
        return ops.NOQ
    d = torch.rand(max(n, 1)) * 0.02 + 0.01
    z = torch.round(torch.rand(max(n, 1)) * 200)
    if mode == "scalar":
        return ops.qparam_from_ckpt(d[0], z[0], 255.0, dev)
    return ops.qparam_from_ckpt(d.view(1, 1, -1), z.view(1, 1, -1), 255.0, dev)


only = sys.argv[1] if len(sys.argv) > 1 else ""
for m, f, k in [(16384, 5120, 1280), (65536, 2560, 640)]:
    a = torch.randn(m, k, device=dev).half()
    b = torch.randint(-15, 16, (2 * f, k), device=dev).half()
    scale = torch.rand(2 * f, device=dev) * 0.01
    bias = torch.rand(2 * f, device=dev)
    fl = 2.0 * m * 2 * f * k
    if not only:
        t = timeit(lambda: ops.gemm(a, b, 2 * f, scale=scale, bias=bias, want_f32=True))
        print(f"geglu-shape {m}x{2*f}x{k} plain f32 out: {t:.3f} ms {fl/t/1e9:.0f} TF", flush=True)
        g32 = ops.gemm(a, b, 2 * f, scale=scale, bias=bias, want_f32=True)
        q2 = qp(f, "kwise")
        t = timeit(lambda: ops.geglu_quant(g32, q2))
        print(f"   geglu_quant kernel: {t:.3f} ms", flush=True)
        del g32
    for mode in ["none", "scalar", "kwise"]:
        q2 = qp(f, mode)
        t = timeit(lambda: ops.gemm(a, b, 2 * f, scale=scale, bias=bias, epi=ops.EPI_GEGLU, q2=q2))
        print(f"   fused GEGLU q2={mode}: {t:.3f} ms {fl/t/1e9:.0f} TF", flush=True)
for bsz, t_, heads, d, k in [(16, 1024, 20, 64, 1280), (16, 4096, 10, 64, 640)]:
    m, n = bsz * t_, heads * d
    a = torch.randn(m, k, device=dev).half()
    w = torch.randint(-15, 16, (n, k), device=dev).half()
    scale = torch.rand(n, device=dev) * 0.01
    fl = 2.0 * m * n * k
    if not only:
        t = timeit(lambda: ops.gemm(a, w, n, scale=scale, want_f32=True))
        print(f"qkv-shape {m}x{n}x{k} plain f32 out: {t:.3f} ms {fl/t/1e9:.0f} TF", flush=True)
        g32 = ops.gemm(a, w, n, scale=scale, want_f32=True)
        for tr in (False, True):
            q2 = qp(d, "kwise")
            t = timeit(lambda: ops.qkv_pack(g32, bsz, t_, heads, d, 64, transpose=tr, q=q2))
            print(f"   qkv_pack kernel transpose={tr}: {t:.3f} ms", flush=True)
    for mode in ["none", "kwise"]:
        for tr in (False, True):
            q2 = qp(d, mode)
            dst = ops.qkv_dest(bsz, t_, heads, d, 64, tr, dev)
            t = timeit(lambda: ops.gemm(a, w, n, scale=scale, epi=ops.EPI_QKV, q2=q2, out=dst,
                                        qkv=(heads, d, 64, t_, t_, tr, False)))
            print(f"   fused QKV q2={mode} transpose={tr}: {t:.3f} ms {fl/t/1e9:.0f} TF", flush=True)
