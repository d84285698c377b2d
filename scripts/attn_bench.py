"""Time dgq_attention (both passes) on the SDXL / SD attention shapes; CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgq_b200 import ops

SHAPES = [  # (b, heads, t, s, d, start_peak, label)
    (16, 10, 4096, 4096, 64, False, "sdxl self 64x64"),
    (16, 20, 1024, 1024, 64, False, "sdxl self 32x32"),
    (16, 10, 4096, 77, 64, True, "sdxl cross 64x64"),
    (16, 20, 1024, 77, 64, True, "sdxl cross 32x32"),
    (16, 8, 4096, 4096, 40, False, "sd self 64x64"),
]


def main():
    one = "--one" in sys.argv
    dev = "cuda"
    res = []
    for b, h, t, s, d, sp, label in (SHAPES[1:2] if one else SHAPES):
        dp = (d + 63) // 64 * 64
        x = torch.randn(b * t, h * d, device=dev)
        kx = torch.randn(b * s, h * d, device=dev)
        q = ops.qkv_pack(x, b, t, h, d, dp)
        k = ops.qkv_pack(kx, b, s, h, d, dp)
        v = ops.qkv_pack(kx, b, s, h, d, dp, transpose=True)
        out = torch.empty(b * t, h * d, device=dev)
        n = 1 if one else 5
        for _ in range(1 if one else 2):
            ops.attention(q, k, v, d, map_mode=ops.MAP_LOG2, real_time=True, start_peak=sp, out=out)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            ops.attention(q, k, v, d, map_mode=ops.MAP_LOG2, real_time=True, start_peak=sp, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = 4.0 * b * h * t * s * d
        res.append(dict(label=label, ms=ms, useful_tflops=flops / ms / 1e9))
        print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/attn_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
