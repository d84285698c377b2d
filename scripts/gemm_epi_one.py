"""one launch of the K = 128 (epilogue-only) qGEMM of each kind, for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgq_b200 import ops
from scripts.gemm_i8_bench import operands
dev = "cuda"
m, n, k = 16384, 10240, 128
a8, b8, colsum, b_off, az, ad, a16, b16, scale = operands(m, n, k, 4, dev)
out = torch.empty(m, n, dtype=torch.float16, device=dev)
for _ in range(2):
    ops.gemm(a8, b8, n, scale=scale, out=out, row_scale=ad, row_zp=az, colsum=colsum, b_off=b_off)
    ops.gemm(a16, b16, n, scale=scale, out=out, row_scale=ad)
torch.cuda.synchronize()
