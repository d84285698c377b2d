"""Timeline of the pass-2 attention roles on CTA 0: the SM clock at which, per score-tile step, the
QK^T issuer reached the step / sent its first MMA / issued the S-full commit, the PV issuer sent P'V, and the softmax
group asked for S, got it and published P'.  Needs the tracing build of attention.cu:

    python scripts/attn_trace.py --build          # here (nvcc): dgq_b200/_C/libdgq_b200_trace.so, -DDGQ_ATTN_TRACE
    python scripts/attn_trace.py [shape index]    # on the GPU box (shapes of scripts/attn_bench.py)
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dgq_b200._lib as L  # noqa: E402
from dgq_b200 import build as B  # noqa: E402

TRACE_LIB = os.path.join(B.OUT_DIR, "libdgq_b200_trace.so")

if "--build" in sys.argv:
    B.build()
    obj = os.path.join(B.OUT_DIR, "attention_trace.o")
    subprocess.run(["nvcc"] + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DDGQ_ATTN_TRACE", "-c", os.path.join(B.CSRC, "attention.cu"), "-o", obj], check=True)
    objs = [os.path.join(B.OUT_DIR, s.replace(".cu", ".o")) for s in B.SOURCES if s != "attention.cu"] + [obj]
    subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", TRACE_LIB] + objs, check=True)
    print(TRACE_LIB)
    sys.exit(0)

L.LIB_PATH = TRACE_LIB
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
names = ["qk_reach", "qk_mma", "qk_commit", "pv_mma", "sm_ask_s", "sm_got_s", "sm_pub_p"]
t0 = tr[1][0]
print(label, "| step |", " ".join(f"{n:>10s}" for n in names))
for u in range(8, 32):
    print(f"{u:4d} |", " ".join(f"{tr[r][u] - t0:10d}" for r in range(7)))


def avg(f, lo=8, hi=56):
    xs = [f(u) for u in range(lo, hi)]
    return sum(xs) / len(xs)


print("cycles per step (same half, u -> u + 2, halved):", avg(lambda u: tr[6][u + 2] - tr[6][u]) / 2)
print("QK^T issuer of a half, per own step: reach -> first MMA", avg(lambda u: tr[1][u] - tr[0][u]), "| MMAs + commit", avg(lambda u: tr[2][u] - tr[1][u]),
      "| commit -> reaches its next step", avg(lambda u: tr[0][u + 2] - tr[2][u]))
print("softmax half: waits for S", avg(lambda u: tr[5][u] - tr[4][u]), "| S -> P' published", avg(lambda u: tr[6][u] - tr[5][u]),
      "| P' published -> PV issued", avg(lambda u: tr[3][u] - tr[6][u]))
