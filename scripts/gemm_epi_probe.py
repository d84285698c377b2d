"""Separates the qGEMM's per-launch cost from its per-tile epilogue cost: K = 128 (one k-block: the MMA time is
negligible) at 1, 5 and 35 tile rounds, kind::i8 and kind::f16, fp16 / fp32+residual results."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dgq_b200 import ops  # noqa: E402
from scripts.gemm_i8_bench import operands, timed  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for m, n, k in [(16384, 256, 128), (16384, 1280, 128), (16384, 10240, 128), (16384, 1280, 1280), (4096, 1280, 1280), (65536, 1280, 1280)]:
    a8, b8, colsum, b_off, az, ad, a16, b16, scale = operands(m, n, k, 4, dev)
    row = f"{m}x{n}x{k}: "
    for out_f32 in (False, True):
        resid = torch.randn(m, n, device=dev) if out_f32 else None
        out = torch.empty(m, n, dtype=torch.float32 if out_f32 else torch.float16, device=dev)
        t8 = timed(lambda: ops.gemm(a8, b8, n, scale=scale, out=out, resid=resid, row_scale=ad, row_zp=az, colsum=colsum, b_off=b_off), flush)
        t16 = timed(lambda: ops.gemm(a16, b16, n, scale=scale, out=out, resid=resid, row_scale=ad), flush)
        row += f"{'f32+resid' if out_f32 else 'f16out'} i8 {t8 * 1e3:.1f} us f16 {t16 * 1e3:.1f} us | "
    print(row, flush=True)
