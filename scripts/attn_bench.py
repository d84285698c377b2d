"""Time dgq_attention (both passes) on the SDXL / SD attention shapes with the headline operand layout (per-channel
aqtizer_q/k/v scales: integer Q, folded hi | lo K) and the to_out quantizer fused; CUDA events.

    python scripts/attn_bench.py            # every shape, split and un-split operands
    python scripts/attn_bench.py --one N    # one launch pair of shape N (for ncu)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dgq_b200 import engine, ops  # noqa: E402

SHAPES = [  # (b, heads, t, s, d, start_peak, label)
    (16, 10, 4096, 4096, 64, False, "sdxl self 64x64"),
    (16, 20, 1024, 1024, 64, False, "sdxl self 32x32"),
    (16, 10, 4096, 77, 64, True, "sdxl cross 64x64"),
    (16, 20, 1024, 77, 64, True, "sdxl cross 32x32"),
    (16, 8, 4096, 4096, 40, False, "sd self 64x64"),
    (16, 8, 1024, 1024, 80, False, "sd self 32x32"),
    (16, 8, 256, 256, 160, False, "sd self 16x16"),
]


def qparam(g, n, dev):
    lab = torch.randint(0, 16, (n,), generator=g)
    lo = -(torch.rand(16, generator=g) * 3 + 1)
    hi = torch.rand(16, generator=g) * 3 + 1
    d = (hi - lo) / 255
    return ops.qparam_from_ckpt(d[lab].view(1, 1, -1), torch.round(-lo / d)[lab].view(1, 1, -1), 255.0, dev)


def main():
    one = int(sys.argv[sys.argv.index("--one") + 1]) if "--one" in sys.argv else None
    dev = "cuda"
    res = []
    for idx, (b, h, t, s, d, sp, label) in enumerate(SHAPES):
        if one is not None and idx != one:
            continue
        g = torch.Generator().manual_seed(idx)
        dp = (d + 63) // 64 * 64
        x = torch.randn(b * t, h * d, generator=g).to(dev)
        kx = torch.randn(b * s, h * d, generator=g).to(dev)
        qq, qk, qv, qo = qparam(g, d, dev), qparam(g, d, dev), qparam(g, d, dev), qparam(g, h * d, dev)
        row = dict(label=label)
        for split in ((True,) if one is not None else (True, False)):
            engine.ATTN_SPLIT = split
            plan = engine.attn_plan(qq, dp)
            q = ops.qkv_pack(x, b, t, h, d, dp, q=qq, emit_int=plan["q_int"])
            k = ops.qkv_pack(kx, b, s, h, d, dp, skip_first=sp, q=qk, kfold=plan["kfold"], split=plan["split"])
            v = ops.qkv_pack(kx, b, s, h, d, dp, transpose=True, q=qv)
            out = torch.empty(b * t, h * d, dtype=torch.float16, device=dev)
            fn = lambda: ops.attention(q, k, v, d, map_mode=ops.MAP_LOG2, real_time=True, start_peak=sp, out=out,  # noqa: E731
                                       out_q=qo, q_scale=plan["q_scale"], q_period=plan["q_period"], k_split=plan["split"])
            n = 1 if one is not None else 5
            for _ in range(1 if one is not None else 2):
                fn()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            row["split" if split else "rounded"] = dict(ms=round(ms, 4), useful_tflops=round(4.0 * b * h * t * s * d / ms / 1e9, 1))
        engine.ATTN_SPLIT = True
        res.append(row)
        print(row, flush=True)
    if one is None:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open("gpurun_out/attn_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
