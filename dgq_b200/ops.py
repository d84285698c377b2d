"""Thin torch-tensor wrappers over the C ABI (include/dgq_b200.h).

torch is used for device memory and streams only; every computation below is one of the
hand-written sm_100a kernels in dgq_b200/csrc.  Launches go to torch's current CUDA stream, so
they are CUDA-graph capturable.  A module-level launch counter feeds bench.py's `gpu_launches`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import (Q_NONE, Q_SCALAR, Q_KWISE, Q_ROWWISE, MAP_NONE, MAP_UNIFORM, MAP_LOG2,  # noqa: F401
                   EPI_PLAIN, EPI_GEGLU, EPI_QKV)

LAUNCHES = 0  # kernels launched through this module (some entry points launch two)

# dtype of activations BETWEEN kernels (GEMM results, residual stream).  fp32 keeps the inputs of
# every quantizer identical to the reference's to ~1e-6; fp16 halves that traffic but moves ~1 % of
# the values across a rounding boundary.  Tensor-core operands are fp16 either way.
ACT_DTYPE = torch.float16 if __import__("os").environ.get("DGQ_ACT_DTYPE", "fp32") == "fp16" else torch.float32


def _is32(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return 1
    if t.dtype == torch.float16:
        return 0
    raise TypeError(f"activations must be fp16 or fp32, got {t.dtype}")


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


@dataclass
class QParam:
    """A (delta, zero_point) pair resident on the device plus how it is indexed."""
    mode: int = Q_NONE
    delta: Optional[torch.Tensor] = None
    zp: Optional[torch.Tensor] = None
    period: int = 1
    qmax: float = 255.0
    int_ok: bool = False   # |code - zp| <= 2048 for every entry: the integer operand is exact in fp16
    inv: Optional[torch.Tensor] = None   # 1/delta (IEEE), for the quantizer's multiply fast path
    zp_in_range: bool = False   # 0 <= zp <= qmax everywhere: an exact 0 (conv zero padding) is the u8 code zp

    def struct(self, emit_int: int = 0) -> L.QuantT:
        if self.inv is None and self.delta is not None:
            self.inv = torch.reciprocal(self.delta)
        return L.QuantT(_p(self.delta), _p(self.zp), _p(self.inv), self.mode, self.period, self.qmax, int(emit_int))

    @property
    def exact(self) -> bool:
        """scalar / row-wise scales can leave the GEMM exact: integer A, delta applied per row in the
        epilogue.  K-wise scales vary along the reduction and must be folded into the fp16 operand."""
        return self.int_ok and self.mode in (Q_SCALAR, Q_ROWWISE)


NOQ = QParam()


def qparam_from_ckpt(delta: torch.Tensor, zp: torch.Tensor, qmax: float, device, *, conv: bool = False,
                     kperm=None) -> QParam:
    """Map a checkpoint (delta, zp) of shape (), (1,1,X) or (1,X,1) to a device QParam
    (SURVEY.md 8a').  Linear inputs are (B,T,C): the last axis is K, the middle axis is the row.
    `conv`: the quantizer saw the unfolded (B, C*kh*kw, L) tensor, so the axes swap.  `kperm`
    re-orders a K-wise table into the GEMM's K order (conv: tap-major)."""
    d, z = _f32(delta, device), _f32(zp, device)
    int_ok = bool((z.abs().max() + qmax <= 2048).item())
    in_range = bool(((z.min() >= 0) & (z.max() <= qmax)).item())
    if d.dim() == 0 or d.numel() == 1 and d.dim() <= 1:
        d1 = d.reshape(1)
        return QParam(Q_SCALAR, d1, z.reshape(1).expand(1).contiguous(), 1, qmax, int_ok, torch.reciprocal(d1), in_range)
    if d.dim() == 3 and d.shape[0] == 1 and d.shape[1] == 1:      # (1,1,X): last axis
        mode = Q_ROWWISE if conv else Q_KWISE
    elif d.dim() == 3 and d.shape[0] == 1 and d.shape[2] == 1:    # (1,X,1): middle axis
        mode = Q_KWISE if conv else Q_ROWWISE
    else:
        raise ValueError(f"unsupported quantizer parameter shape {tuple(d.shape)}")
    d, z = d.reshape(-1), z.reshape(-1).expand(d.numel())
    if kperm is not None and mode == Q_KWISE:
        d, z = d[kperm], z[kperm]
    d = d.contiguous()
    return QParam(mode, d, z.contiguous(), d.numel(), qmax, int_ok, torch.reciprocal(d), in_range)


# ------------------------------------------------------------------------------------------
def fake_quant(x: torch.Tensor, delta: torch.Tensor, zp: torch.Tensor, period: int, inner: int, qmax: float,
               want_codes: bool = False):
    """UniformAffineQuantizer.forward on an fp32 tensor; index = (i // inner) % period."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty_like(x)
    codes = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_codes else None
    L.check(L.lib().dgq_fake_quant_f32(_p(x), x.numel(), _p(delta), _p(zp), period, inner, qmax, _p(out),
                                       _p(codes), _stream()), "dgq_fake_quant_f32")
    _count()
    return (out, codes) if want_codes else out


def max_f32(x: torch.Tensor) -> torch.Tensor:
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    scratch = torch.empty(1024, dtype=torch.float32, device=x.device)
    L.check(L.lib().dgq_max_f32(_p(x), x.numel(), _p(out), _p(scratch), _stream()), "dgq_max_f32")
    _count(2)
    return out


def t2i_log_quant(x: torch.Tensor, delta: Optional[torch.Tensor], qmax: float, want_codes: bool = False):
    """T2ILogQuantizer.forward; delta None => real-time (x.max())."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    d = max_f32(x) if delta is None else delta
    out = torch.empty_like(x)
    codes = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_codes else None
    L.check(L.lib().dgq_t2i_log_quant_f32(_p(x), x.numel(), _p(d), qmax, _p(out), _p(codes), _stream()),
            "dgq_t2i_log_quant_f32")
    _count()
    return (out, codes) if want_codes else out


def pack_weight(w: torch.Tensor, delta, zp, alpha, qmax: float, use_wq: bool, *, ci_pad: Optional[int] = None,
                n_pad: Optional[int] = None, want_codes: bool = False, want_packed4: bool = False):
    """w fp32 [n, ci, kh, kw] or [n, k] -> fp16 operand [n_pad, taps*ci_pad] (K order tap-major)."""
    assert w.is_cuda and w.dtype == torch.float32
    w = w.contiguous()
    n, ci = w.shape[0], w.shape[1]
    taps = w[0, 0].numel() if w.dim() == 4 else 1
    ci_pad = ci_pad or (ci + 7) // 8 * 8
    n_pad = n_pad or (n + 7) // 8 * 8
    k_out = taps * ci_pad
    operand = torch.empty(n_pad, k_out, dtype=torch.float16, device=w.device)
    codes = torch.empty(n_pad, k_out, dtype=torch.uint8, device=w.device) if (want_codes or want_packed4) else None
    packed = torch.empty(n_pad, k_out // 2, dtype=torch.uint8, device=w.device) if want_packed4 else None
    d = _f32(delta, w.device).reshape(-1) if delta is not None else None
    z = _f32(zp, w.device).reshape(-1) if zp is not None else None
    if z is not None and z.numel() == 1 and n > 1:
        z = z.expand(n).contiguous()
    a = _f32(alpha, w.device) if alpha is not None else None
    L.check(L.lib().dgq_pack_weight(_p(w), _p(d), _p(z), _p(a), n, ci, taps, ci_pad, n_pad, qmax, int(use_wq),
                                    _p(codes), _p(packed), _p(operand), _stream()), "dgq_pack_weight")
    _count(2 if want_packed4 else 1)
    return operand, codes, packed


def unpack_weight(codes: torch.Tensor, bits: int, zp: torch.Tensor, n: int, ci: int, taps: int, ci_pad: int,
                  n_pad: int) -> torch.Tensor:
    """compiled-checkpoint codes (uint8; two per byte when bits == 4) -> resident fp16 operand (code - zp)."""
    assert codes.is_cuda and codes.dtype == torch.uint8 and codes.is_contiguous()
    operand = torch.empty(n_pad, taps * ci_pad, dtype=torch.float16, device=codes.device)
    z = _f32(zp, codes.device).reshape(-1)
    L.check(L.lib().dgq_unpack_weight(_p(codes), bits, _p(z), n, ci, taps, ci_pad, n_pad, _p(operand), _stream()),
            "dgq_unpack_weight")
    _count()
    return operand


def act_producer(src0: torch.Tensor, *, batch: int, h: int, w: int, src1: Optional[torch.Tensor] = None,
                 upsample: bool = False, ksize: int = 1, stride: int = 1, gn=None, act: int = 0,
                 q: QParam = NOQ, pad_quantized: bool = False, ldo: Optional[int] = None,
                 want_codes: bool = False, emit_int: int = 0):
    """src NHWC ([batch, hs, ws, c], fp16 or fp32) -> fp16 A operand [M, ldo] (+ codes); emit_int = 2: the u8
    code operand of the kind::i8 GEMM.  gn = (mean, rstd, gamma, beta) or None."""
    c0 = src0.shape[-1]
    c1 = src1.shape[-1] if src1 is not None else 0
    pad = 1 if ksize == 3 else 0
    ho = (h + 2 * pad - ksize) // stride + 1
    wo = (w + 2 * pad - ksize) // stride + 1
    K = ksize * ksize * (c0 + c1)
    ldo = ldo or K
    M = batch * ho * wo
    emit_int = int(emit_int)
    out = torch.empty(M, ldo, dtype=torch.uint8 if emit_int == 2 else torch.float16, device=src0.device)
    codes = torch.empty(M, K, dtype=torch.uint8, device=src0.device) if want_codes else None
    a = L.ProducerT(_p(src0), _p(src1), c0, c1, int(src0.dtype == torch.float32), batch, h, w, int(upsample),
                    ksize, stride, pad,
                    _p(gn[0]) if gn else None, _p(gn[1]) if gn else None, _p(gn[2]) if gn else None,
                    _p(gn[3]) if gn else None, act, q.struct(emit_int), int(pad_quantized), _p(out), ldo, _p(codes))
    L.check(L.lib().dgq_act_producer(C.byref(a), _stream()), "dgq_act_producer")
    _count()
    return (out, codes) if want_codes else out


def gn_stats(src0: torch.Tensor, src1: Optional[torch.Tensor], batch: int, hw: int, eps: float):
    c0 = src0.shape[-1]
    c1 = src1.shape[-1] if src1 is not None else 0
    dev = src0.device
    mean = torch.empty(batch, 32, dtype=torch.float32, device=dev)
    rstd = torch.empty(batch, 32, dtype=torch.float32, device=dev)
    scratch = torch.empty(batch * 64 * 64, dtype=torch.float32, device=dev)
    L.check(L.lib().dgq_gn_stats(_p(src0), _p(src1), _is32(src0), c0, c1, batch, hw, eps, _p(mean), _p(rstd),
                                 _p(scratch), _stream()), "dgq_gn_stats")
    _count(2)
    return mean, rstd


def _emit_modes(qs, emit_int):
    """per-quantizer operand form: 0 de-quantised fp16, 1 integer (code - zp) fp16, 2 u8 codes.  A scalar applies to
    every quantizer that allows it (K-wise scales must be folded: always 0)."""
    if isinstance(emit_int, (list, tuple)):
        return [int(e) if q.exact else 0 for e, q in zip(emit_int, qs)]
    return [int(emit_int) if q.exact else 0 for q in qs]


def _row_outputs(x, qs, emit_int=0):
    m, c = x.shape
    modes = _emit_modes(qs, emit_int)
    outs = [torch.empty(m, c, dtype=torch.uint8 if e == 2 else torch.float16, device=x.device) for e in modes]
    qarr = (L.QuantT * len(qs))(*[q.struct(e) for q, e in zip(qs, modes)])
    oarr = (C.c_void_p * len(qs))(*[o.data_ptr() for o in outs])
    return outs, qarr, oarr


def ln_quant(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, qs: Sequence[QParam],
             emit_int=0):
    """x [m, c] -> [fp16 [m, c]] * len(qs); emit_int: integer operands for the quantizers that allow it."""
    outs, qarr, oarr = _row_outputs(x, qs, emit_int)
    L.check(L.lib().dgq_ln_quant(_p(x), _is32(x), x.shape[0], x.shape[1], _p(gamma), _p(beta), eps, len(qs), qarr,
                                 oarr, _stream()), "dgq_ln_quant")
    _count()
    return outs


def row_quant(x: torch.Tensor, qs: Sequence[QParam], want_codes: bool = False, emit_int=0):
    outs, qarr, oarr = _row_outputs(x, qs, emit_int)
    codes = [torch.empty(x.shape, dtype=torch.uint8, device=x.device) for _ in qs] if want_codes else None
    carr = (C.c_void_p * len(qs))(*[c.data_ptr() for c in codes]) if want_codes else None
    L.check(L.lib().dgq_row_quant(_p(x), int(x.dtype == torch.float32), x.shape[0], x.shape[1], len(qs), qarr,
                                  oarr, carr, _stream()), "dgq_row_quant")
    _count()
    return (outs, codes) if want_codes else outs


def geglu_quant(x: torch.Tensor, q: QParam, emit_int: int = 0) -> torch.Tensor:
    m, f2 = x.shape
    e = _emit_modes([q], emit_int)[0]
    out = torch.empty(m, f2 // 2, dtype=torch.uint8 if e == 2 else torch.float16, device=x.device)
    L.check(L.lib().dgq_geglu_quant(_p(x), _is32(x), m, f2 // 2, q.struct(e), _p(out), _stream()),
            "dgq_geglu_quant")
    _count()
    return out


def gemm(a: torch.Tensor, b: torch.Tensor, n: int, *, scale=None, bias=None, temb=None, rows_per_batch: int = 0,
         resid=None, out: Optional[torch.Tensor] = None, want_f32: bool = False, k: Optional[int] = None,
         row_scale: Optional[torch.Tensor] = None, row_period: int = 1, epi: int = L.EPI_PLAIN,
         q2: QParam = NOQ, q2_emit_int: int = 0, qkv: Optional[tuple] = None,
         colsum: Optional[torch.Tensor] = None, b_off: Optional[torch.Tensor] = None,
         row_zp: Optional[torch.Tensor] = None, kfold: Optional[torch.Tensor] = None, k_split: bool = False,
         conv: Optional[tuple] = None):
    """a fp16 [m, lda], b fp16 [n_pad, ldb] -> [m, n] (n multiple of 8), fp32 if want_f32 (or `out`
    is fp32) else fp16.  temb / resid must share one dtype (fp16 or fp32).
    epi = EPI_GEGLU: b rows interleaved (pack_weight geglu=True); returns the fp16 operand [m, n/2] of
    ff.net.2 quantised with q2.  epi = EPI_QKV: `out` is the head-split destination, qkv =
    (heads, d, dp, tokens, tp, transpose, skip_first)."""
    m = a.shape[0]
    k = k or min(a.shape[1], b.shape[1])
    cb = ch = cw = cc = cld = 0
    csoob = None
    if conv is not None:                  # implicit 3x3 conv: `a` is the NHWC u8 code tensor [m = b*h*w, c]
        cb, ch, cw, cc, csoob = conv
        k, cld = 9 * cc, csoob.shape[1]
    i8 = a.dtype == torch.uint8           # u8 activation codes x s8 weight codes: dgq_gemm_i8
    if i8 and (b.dtype != torch.int8 or colsum is None or row_zp is None or row_scale is None):
        raise TypeError("gemm: a u8 A operand needs the s8 weight operand with its colsum / row_zp / row_scale")
    e2 = _emit_modes([q2], q2_emit_int)[0] if q2.mode != Q_NONE else 0
    if epi == L.EPI_QKV and q2.mode != Q_NONE:   # Q operand of the attention kernel: bare integers in EVERY scale layout
        e2 = int(bool(q2_emit_int))
    if epi == L.EPI_GEGLU:
        out = torch.empty(m, n // 2, dtype=torch.uint8 if e2 == 2 else torch.float16, device=a.device)
    elif epi == L.EPI_QKV:
        assert out is not None and out.dtype == torch.float16 and qkv is not None
    elif out is None:
        out = torch.empty(m, n, dtype=torch.float32 if want_f32 else torch.float16, device=a.device)
    o32 = out.dtype == torch.float32
    ep32 = 0
    for e in (temb, resid):
        if e is not None:
            ep32 = _is32(e)
    if temb is not None and resid is not None and _is32(temb) != _is32(resid):
        raise TypeError("temb and resid must have the same dtype")
    heads, d, dp, tokens, tp, transpose, skip_first = qkv if qkv is not None else (0, 0, 0, 0, 0, 0, 0)
    ldc = out.stride(0) if epi != L.EPI_QKV else 8
    g = L.GemmT(_p(a), a.stride(0), _p(b), b.stride(0), m, n, k, _p(scale), _p(row_scale), row_period, _p(bias),
                _p(temb), rows_per_batch,
                temb.stride(0) if temb is not None else 0, _p(resid), resid.stride(0) if resid is not None else 0,
                None if o32 else _p(out), ldc, _p(out) if o32 else None, ep32,
                epi, q2.struct(e2), heads, d, dp, tokens, tp, int(transpose), int(skip_first),
                _p(kfold), int(k_split), _p(colsum), _p(b_off), _p(row_zp), cb, ch, cw, cc, _p(csoob), cld)
    if i8:
        L.check(L.lib().dgq_gemm_i8(C.byref(g), _stream()), "dgq_gemm_i8")
    else:
        L.check(L.lib().dgq_gemm_f16(C.byref(g), _stream()), "dgq_gemm_f16")
    _count()
    return out


def conv_oob_colsum(operand: torch.Tensor, c: int) -> torch.Tensor:
    """s8 conv operand [n_pad, 9*c] (tap-major) -> int32 [9, n_pad]: per border class, the column sums of the taps
    that fall outside the image (dgq_gemm_i8's implicit-conv zero-padding correction)."""
    n_pad = operand.shape[0]
    out = torch.empty(9, n_pad, dtype=torch.int32, device=operand.device)
    L.check(L.lib().dgq_conv_oob_colsum(_p(operand), n_pad, c, _p(out), _stream()), "dgq_conv_oob_colsum")
    _count()
    return out


def weight_to_i8(codes: torch.Tensor, zp: torch.Tensor, n: int, qmax: float):
    """u8 weight codes [n_pad, k] (dgq_pack_weight) -> (s8 operand, colsum int32 [n_pad], b_off int32 [n_pad] | None)."""
    n_pad, k = codes.shape
    operand = torch.empty(n_pad, k, dtype=torch.int8, device=codes.device)
    colsum = torch.empty(n_pad, dtype=torch.int32, device=codes.device)
    b_off = torch.empty(n_pad, dtype=torch.int32, device=codes.device) if qmax > 127 else None
    z = _f32(zp, codes.device).reshape(-1)
    if z.numel() == 1 and n > 1:
        z = z.expand(n).contiguous()
    L.check(L.lib().dgq_weight_to_i8(_p(codes), _p(z), n, n_pad, k, qmax, _p(operand), _p(colsum), _p(b_off), _stream()),
            "dgq_weight_to_i8")
    _count()
    return operand, colsum, b_off


def qkv_dest(b: int, t: int, heads: int, d: int, dp: int, transpose: bool, device, split: bool = False) -> torch.Tensor:
    """destination of an EPI_QKV GEMM (zero-filled only when it has padding the kernel does not write);
    split: the K operand as an fp16 (hi | lo) pair, [b, heads, t, 2 dp]"""
    tp = (t + 7) // 8 * 8
    shape = (b, heads, dp, tp) if transpose else (b, heads, t, 2 * dp if split else dp)
    padded = dp != d or (transpose and tp != t)
    return (torch.zeros if padded else torch.empty)(shape, dtype=torch.float16, device=device)


def qkv_pack(x: torch.Tensor, b: int, t: int, heads: int, d: int, dp: int, *, transpose: bool = False,
             skip_first: bool = False, q: QParam = NOQ, emit_int: bool = False, kfold: Optional[torch.Tensor] = None,
             split: bool = False) -> torch.Tensor:
    """emit_int: bare integers code - zp (the Q operand); kfold / split: the K operand, value * kfold[channel] as an
    fp16 (hi | lo) pair [b, heads, t, 2 dp]"""
    tp = (t + 7) // 8 * 8
    shape = (b, heads, dp, tp) if transpose else (b, heads, t, 2 * dp if split else dp)
    out = torch.empty(shape, dtype=torch.float16, device=x.device)
    L.check(L.lib().dgq_qkv_pack(_p(x), _is32(x), x.stride(-2), b, t, heads, d, dp, tp, int(transpose),
                                 int(skip_first), q.struct(int(bool(emit_int)) if q.mode != Q_NONE else 0), _p(kfold),
                                 int(split), _p(out), _stream()), "dgq_qkv_pack")
    _count()
    return out


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, d: int, *, map_mode: int, real_time: bool = False,
              start_peak: bool = False, delta: Optional[torch.Tensor] = None, qmax: float = 255.0,
              out: Optional[torch.Tensor] = None, want_codes: bool = False, out_dtype=torch.float16,
              out_q: Optional[QParam] = None, out_emit_int: int = 0, q_scale: Optional[torch.Tensor] = None,
              q_period: int = 1, k_split: bool = False):
    """q [b,h,t,dp], k [b,h,s,dp] (k_split: [b,h,s,2dp] hi | lo), vt [b,h,dp,sp] fp16 -> [b*t, h*d];
    q_scale[token % q_period]: delta of an integer Q operand.  Returns (out, rt_delta[, codes])."""
    b, heads, t, dp = q.shape
    s, sp = k.shape[2], vt.shape[3]
    dev = q.device
    oe = _emit_modes([out_q], out_emit_int)[0] if out_q is not None and out_q.mode != Q_NONE else 0
    if out_q is not None:     # fused quantizer of the consumer: the result IS its GEMM operand (fp16, or u8 codes)
        out_dtype = torch.uint8 if oe == 2 else torch.float16
    if out is None:
        out = torch.empty(b * t, heads * d, dtype=out_dtype, device=dev)
    row_max = torch.empty(b * heads * t, dtype=torch.float32, device=dev)
    row_sum = torch.empty(b * heads * t, dtype=torch.float32, device=dev)
    # gmax[0]: real-time delta, zeroed by dgq_attention itself (a memset node, not an extra fill kernel)
    gmax = torch.empty(1, dtype=torch.float32, device=dev) if real_time else torch.zeros(1, dtype=torch.float32, device=dev)
    codes = torch.zeros(b, heads, t, s, dtype=torch.uint8, device=dev) if want_codes else None
    a = L.AttnT(_p(q), _p(k), _p(vt), b, heads, t, s, sp, d, dp, float(d) ** -0.5, map_mode, int(real_time),
                int(start_peak), _p(delta), qmax, _p(row_max), _p(row_sum), _p(gmax), _p(out), out.stride(0),
                int(out.dtype == torch.float32), _p(codes), (out_q or NOQ).struct(oe), _p(q_scale), int(q_period),
                int(k_split))
    L.check(L.lib().dgq_attention(C.byref(a), _stream()), "dgq_attention")
    _count(2)
    return (out, gmax[:1], codes) if want_codes else (out, gmax[:1])


def timestep_embedding(t: torch.Tensor, dim: int, *, f32: bool = False) -> torch.Tensor:
    n = t.numel()
    t = t.to(torch.float32).contiguous()
    out = torch.empty(n, dim, dtype=torch.float32 if f32 else torch.float16, device=t.device)
    L.check(L.lib().dgq_timestep_embedding(_p(t), n, dim, None if f32 else _p(out), _p(out) if f32 else None, dim,
                                           _stream()), "dgq_timestep_embedding")
    _count()
    return out


def nchw_to_nhwc(x: torch.Tensor, c_pad: int, dtype=None) -> torch.Tensor:
    b, c, h, w = x.shape
    out = torch.empty(b, h, w, c_pad, dtype=dtype or ACT_DTYPE, device=x.device)
    x = x.contiguous()
    L.check(L.lib().dgq_nchw_to_nhwc(_p(x), b, c, h * w, c_pad, _p(out), _is32(out), _stream()), "dgq_nchw_to_nhwc")
    _count()
    return out


def nhwc_to_nchw(x: torch.Tensor, b: int, c: int, h: int, w: int) -> torch.Tensor:
    out = torch.empty(b, c, h, w, dtype=torch.float32, device=x.device)
    L.check(L.lib().dgq_nhwc_to_nchw(_p(x), _is32(x), b, c, h * w, x.stride(-2), _p(out), _stream()),
            "dgq_nhwc_to_nchw")
    _count()
    return out


def silu(x: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(x)
    L.check(L.lib().dgq_silu(_p(x), _is32(x), x.numel(), _p(out), _stream()), "dgq_silu")
    _count()
    return out


def softmax_rows(s: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> fp16 softmax(scale * s) along the rows (VAE mid-block attention map)"""
    if s.dtype != torch.float32 or s.dim() != 2 or s.stride(1) != 1:
        raise TypeError("softmax_rows: fp32 row-major matrix expected")
    if out is None:
        out = torch.empty(s.shape, dtype=torch.float16, device=s.device)
    L.check(L.lib().dgq_softmax_rows(_p(s), s.shape[0], s.shape[1], s.stride(0), float(scale), _p(out), out.stride(0),
                                     _stream()), "dgq_softmax_rows")
    _count()
    return out


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(a)
    if a.dtype != b.dtype:
        raise TypeError("add: dtype mismatch")
    L.check(L.lib().dgq_add(_p(a), _p(b), _is32(a), a.numel(), _p(out), _stream()), "dgq_add")
    _count()
    return out
