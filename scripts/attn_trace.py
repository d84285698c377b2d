"""Timeline of the pass-2 attention roles on CTA 0 (needs the tracing build of attention.cu: see DESIGN.md 4.2):
per score-tile step, the SM clock at which the QK^T issuer issued / committed, the PV issuer issued / committed and the
softmax group began waiting for S, got S, released S, waited for / got the P' buffer and published P'."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dgq_b200._lib as L  # noqa: E402
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), "libdgq_b200_trace.so")
import torch  # noqa: E402
from dgq_b200 import engine, ops  # noqa: E402
from scripts.attn_bench import SHAPES, qparam  # noqa: E402

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
b, h, t, s, d, sp, label = SHAPES[idx]
dev = "cuda"
g = torch.Generator().manual_seed(idx)
dp = (d + 63) // 64 * 64
x = torch.randn(b * t, h * d, generator=g).to(dev)
kx = torch.randn(b * s, h * d, generator=g).to(dev)
qq, qk, qv, qo = qparam(g, d, dev), qparam(g, d, dev), qparam(g, d, dev), qparam(g, h * d, dev)
plan = engine.attn_plan(qq, dp)
q = ops.qkv_pack(x, b, t, h, d, dp, q=qq, emit_int=plan["q_int"])
k = ops.qkv_pack(kx, b, s, h, d, dp, skip_first=sp, q=qk, kfold=plan["kfold"], split=plan["split"])
v = ops.qkv_pack(kx, b, s, h, d, dp, transpose=True, q=qv)
out = torch.empty(b * t, h * d, dtype=torch.float16, device=dev)
for _ in range(2):
    ops.attention(q, k, v, d, map_mode=ops.MAP_LOG2, real_time=True, start_peak=sp, out=out, out_q=qo,
                  q_scale=plan["q_scale"], q_period=plan["q_period"], k_split=plan["split"])
torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 512))()
lib = L.lib()
lib.dgq_attn_trace_dump.argtypes = [C.c_void_p]
assert lib.dgq_attn_trace_dump(buf) == 0
tr = [[buf[r * 512 + i] for i in range(512)] for r in range(8)]
t0 = tr[2][0]
print(label, "QK issuer, per step uu: loop top | after S-empty wait | MMAs issued | S-full commit issued | all commits issued | after syncwarp   (K tile g = uu // 2: asks | has)")
for uu in range(8, 40):
    g_ = uu // 2
    print(f"{uu:4d} | {tr[6][uu]-t0:8d} {tr[2][uu]-t0:8d} {tr[3][uu]-t0:8d} {tr[4][uu]-t0:8d} {tr[5][uu]-t0:8d} {tr[7][uu]-t0:8d}   | K {tr[0][g_]-t0:8d} {tr[1][g_]-t0:8d}")
