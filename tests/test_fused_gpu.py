"""Fused epilogues (GEGLU / head-split+quantize in the qGEMM, to_out quantizer in the attention
kernel) against the same work done by the stand-alone kernels: results must be IDENTICAL bits --
same accumulators, same formulas, and the reciprocal-path quantizer returns the same codes as the
IEEE-division one (quant/quant_layer.py:295-299)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _q(ops, g, n, mode, level=256):
    if mode == "none":
        return ops.NOQ
    lab = torch.randint(0, 8, (max(n, 1),), generator=g)
    lo = -(torch.rand(8, generator=g) * 3 + 1)
    hi = torch.rand(8, generator=g) * 3 + 1
    lo[0], hi[0] = 0.5, 2.0
    d = (hi - lo) / (level - 1)
    z = torch.round(-lo / d)
    if mode == "scalar":
        return ops.qparam_from_ckpt(d[3], z[3], level - 1.0, DEV)
    view = (1, 1, -1) if mode == "kwise" else (1, -1, 1)
    return ops.qparam_from_ckpt(d[lab].view(view), z[lab].view(view), level - 1.0, DEV)


def _interleave(n):
    i = torch.arange(n)
    return (i // 64) * 32 + i % 32 + ((i % 64) >= 32) * (n // 2)


@pytest.mark.parametrize("m,f,k", [(256, 64, 128), (300, 320, 192), (4096, 1280, 320), (16384, 5120, 640)])
@pytest.mark.parametrize("mode,emit", [("kwise", False), ("scalar", True), ("rowwise", True), ("none", False),
                                       ("scalar", 2), ("rowwise", 2)])     # 2: u8 codes for a kind::i8 consumer
def test_geglu_epilogue(m, f, k, mode, emit):
    from dgq_b200 import ops
    g = torch.Generator().manual_seed(m + f + k)
    a = (torch.randn(m, k, generator=g) * 0.5).half().to(DEV)
    b = torch.randint(-15, 16, (2 * f, k), generator=g).half()
    scale = (torch.rand(2 * f, generator=g) * 0.01 + 0.002)
    bias = torch.randn(2 * f, generator=g) * 0.2
    rows = 64 if m % 64 == 0 else m
    q2 = _q(ops, g, f if mode == "kwise" else rows, mode)
    plain = ops.gemm(a, b.to(DEV), 2 * f, scale=scale.to(DEV), bias=bias.to(DEV), want_f32=True)
    ref = ops.geglu_quant(plain, q2, emit_int=emit)
    perm = _interleave(2 * f)
    out = ops.gemm(a, b[perm].contiguous().to(DEV), 2 * f, scale=scale[perm].to(DEV), bias=bias[perm].to(DEV),
                   epi=ops.EPI_GEGLU, q2=q2, q2_emit_int=emit)
    assert out.shape == ref.shape
    assert torch.equal(out, ref), (out.float() - ref.float()).abs().max().item()


@pytest.mark.parametrize("b_,t,heads,d", [(2, 64, 8, 40), (2, 77, 10, 64), (3, 256, 8, 160), (16, 1024, 20, 64)])
@pytest.mark.parametrize("mode", ["kwise", "scalar", "rowwise", "none"])
@pytest.mark.parametrize("transpose,skip", [(False, False), (False, True), (True, False)])
def test_qkv_epilogue(b_, t, heads, d, mode, transpose, skip):
    from dgq_b200 import ops
    g = torch.Generator().manual_seed(t + heads + d)
    k = 320
    n = heads * d
    m = b_ * t
    dp = (d + 63) // 64 * 64
    a = (torch.randn(m, k, generator=g) * 0.5).half().to(DEV)
    w = torch.randint(-15, 16, (n, k), generator=g).half().to(DEV)
    scale = (torch.rand(n, generator=g) * 0.01 + 0.002).to(DEV)
    q2 = _q(ops, g, d if mode == "kwise" else t - int(skip), mode)
    plain = ops.gemm(a, w, n, scale=scale, want_f32=True)
    ref = ops.qkv_pack(plain, b_, t, heads, d, dp, transpose=transpose, skip_first=skip, q=q2)
    dst = ops.qkv_dest(b_, t, heads, d, dp, transpose, DEV)
    ops.gemm(a, w, n, scale=scale, epi=ops.EPI_QKV, q2=q2, out=dst,
             qkv=(heads, d, dp, t, (t + 7) // 8 * 8, transpose, skip))
    assert dst.shape == ref.shape
    assert torch.equal(dst, ref), (dst.float() - ref.float()).abs().max().item()


@pytest.mark.parametrize("mode,emit", [("kwise", False), ("scalar", True), ("rowwise", True), ("scalar", 2), ("rowwise", 2)])
@pytest.mark.parametrize("t,s,heads,d", [(256, 256, 8, 40), (1024, 77, 10, 64), (4096, 4096, 2, 64)])
def test_attention_out_quant(t, s, heads, d, mode, emit):
    from dgq_b200 import ops
    g = torch.Generator().manual_seed(t + s + d)
    b_ = 2
    dp = (d + 63) // 64 * 64
    q = ops.qkv_pack((torch.randn(b_ * t, heads * d, generator=g)).to(DEV), b_, t, heads, d, dp)
    k = ops.qkv_pack((torch.randn(b_ * s, heads * d, generator=g)).to(DEV), b_, s, heads, d, dp)
    v = ops.qkv_pack((torch.randn(b_ * s, heads * d, generator=g)).to(DEV), b_, s, heads, d, dp, transpose=True)
    oq = _q(ops, g, heads * d if mode == "kwise" else t, mode)
    o32, _ = ops.attention(q, k, v, d, map_mode=ops.MAP_LOG2, real_time=True, out_dtype=torch.float32)
    ref = ops.row_quant(o32, [oq], emit_int=emit)[0]
    out, _ = ops.attention(q, k, v, d, map_mode=ops.MAP_LOG2, real_time=True, out_q=oq, out_emit_int=emit)
    assert torch.equal(out, ref), (out.float() - ref.float()).abs().max().item()
