"""Time the fused to_q / to_k / to_v (EPI_QKV) and GEGLU launches of the SDXL 32x32 level (16384 x 1280 x 1280 and
16384 x 10240 x 1280) in both MMA kinds, with the headline operand layout (integer Q, folded hi | lo K, V^T); CUDA events,
L2 flushed between runs.

    python scripts/qkv_bench.py
    python scripts/qkv_bench.py --one     # one launch each of f16 to_q / to_k / to_v and i8 to_q / to_k / to_v (for ncu)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dgq_b200 import ops  # noqa: E402
from scripts.gemm_i8_bench import operands, timed  # noqa: E402


def kwise(g, n, dev):
    d = torch.rand(n, generator=g) * 0.02 + 0.01
    z = torch.round(torch.rand(n, generator=g) * 200)
    return ops.qparam_from_ckpt(d.view(1, 1, -1), z.view(1, 1, -1), 255.0, dev)


def main():
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(0)
    res = {}
    one = "--one" in sys.argv
    for (b, t, heads, d) in [(16, 1024, 20, 64)] if one else [(16, 1024, 20, 64), (16, 4096, 10, 64)]:
        m, n, k = b * t, heads * d, heads * d
        a8, b8, colsum, b_off, az, ad, a16, b16, scale = operands(m, n, k, 4, dev)
        q2 = kwise(g, d, dev)
        kf = (torch.rand(d, generator=g) * 0.02 + 0.01).to(dev)
        for kind, A, B, kw in (("f16", a16, b16, dict(row_scale=ad)),
                               ("i8", a8, b8, dict(row_scale=ad, row_zp=az, colsum=colsum, b_off=b_off))):
            for name, tr, extra in (("q", False, dict(q2_emit_int=1)), ("k", False, dict(kfold=kf, k_split=True)),
                                    ("v", True, {})):
                dst = ops.qkv_dest(b, t, heads, d, 64, tr, dev, split=(name == "k"))
                fn = lambda: ops.gemm(A, B, n, scale=scale, epi=ops.EPI_QKV, q2=q2, out=dst,  # noqa: E731
                                      qkv=(heads, d, 64, t, t, tr, False), **kw, **extra)
                if one:
                    fn(); torch.cuda.synchronize()
                    continue
                ms = timed(fn, flush)
                res[f"{m}x{n}x{k} {kind} {name}"] = round(ms * 1e3, 1)
                print(f"{m}x{n}x{k} {kind} to_{name}: {ms * 1e3:.1f} us  {2.0 * m * n * k / ms / 1e9:.0f} T/s", flush=True)
    if one:
        return
    m, f, k = 16384, 5120, 1280
    a8, b8, colsum, b_off, az, ad, a16, b16, scale = operands(m, 2 * f, k, 4, dev)
    bias = torch.rand(2 * f, generator=g).to(dev)
    q2 = kwise(g, f, dev)
    qs = ops.qparam_from_ckpt(torch.tensor(0.02), torch.tensor(100.0), 255.0, dev)
    for kind, A, B, kw in (("f16", a16, b16, dict(row_scale=ad)),
                           ("i8", a8, b8, dict(row_scale=ad, row_zp=az, colsum=colsum, b_off=b_off))):
        for qn, q in (("kwise", q2), ("scalar", qs)):
            fn = lambda: ops.gemm(A, B, 2 * f, scale=scale, bias=bias, epi=ops.EPI_GEGLU, q2=q, **kw)  # noqa: E731
            ms = timed(fn, flush)
            res[f"geglu {kind} {qn}"] = round(ms * 1e3, 1)
            print(f"geglu {m}x{2 * f}x{k} {kind} q2={qn}: {ms * 1e3:.1f} us  {2.0 * m * 2 * f * k / ms / 1e9:.0f} T/s", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/qkv_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
