"""Execution of the quantized UNet on dgq_b200's kernels.

The reference runs the UNet module by module in NCHW fp32 with ~8 eager elementwise kernels around
every GEMM.  Here the whole forward stays token-major / NHWC in fp16 and every
op between two GEMMs is ONE fused producer kernel that also applies the activation quantizer of
the consuming QuantLayer:

    GroupNorm stats -> [concat + upsample + GN + SiLU + im2col + quantize]   -> qGEMM(+bias,+temb,+resid)
    [LayerNorm + quantize x3] -> qGEMM x3 -> [head split + quantize] -> fused attention -> ...
    [GEGLU + quantize] -> qGEMM(+resid)

Functions here walk the reference-shaped module tree (duck-typed: QuantLayer, Attention, ...) and
issue kernels through dgq_b200.ops on torch's current stream, so a whole UNet call is capturable
in one CUDA graph.  Module-level `forward`s (API parity) call the same functions after a layout
conversion, so there is exactly one compute path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from . import ops


@dataclass
class Act:
    """activation [b*h*w, c] in ops.ACT_DTYPE (fp32 default); (h, w) = spatial size, or (T, 1) for
    token sequences."""
    t: torch.Tensor
    b: int
    h: int
    w: int

    @property
    def c(self) -> int:
        return self.t.shape[1]

    @property
    def rows(self) -> int:
        return self.h * self.w


# ------------------------------------------------------------------------------------------
# layout shims at the API boundary
# ------------------------------------------------------------------------------------------
def act_from_nchw(x: torch.Tensor, c_pad: Optional[int] = None) -> Act:
    b, c, h, w = x.shape
    c_pad = c_pad or (c + 7) // 8 * 8
    t = ops.nchw_to_nhwc(x.detach().float(), c_pad)
    return Act(t.view(b * h * w, c_pad), b, h, w)


def act_to_nchw(a: Act, c: Optional[int] = None, dtype=torch.float32) -> torch.Tensor:
    c = c or a.c
    return ops.nhwc_to_nchw(a.t, a.b, c, a.h, a.w).to(dtype)


def act_from_tokens(x: torch.Tensor) -> Act:
    b, t, c = x.shape
    return Act(x.detach().reshape(b * t, c).to(ops.ACT_DTYPE).contiguous(), b, t, 1)


def act_to_tokens(a: Act, dtype=torch.float32) -> torch.Tensor:
    return a.t.view(a.b, a.rows, a.c).to(dtype)


# ------------------------------------------------------------------------------------------
# QuantLayer
# ------------------------------------------------------------------------------------------
EXACT_INT = True  # integer A operand + per-row delta in the epilogue wherever the scales allow it
# kind::i8 qGEMM (u8 activation codes x s8 weight codes, exact s32 accumulation, 2x the MMA rate) for every layer
# whose activation scale is constant along K: scalar / row-wise quantizers.  K-wise (DGQ group) scales stay on
# kind::f16 with the scale folded into the operand.  DGQ_I8=0 pins the f16 path (A/B runs).
USE_I8 = os.environ.get("DGQ_I8", "1") != "0"


def _exact(q: ops.QParam) -> bool:
    return EXACT_INT and q.exact


def _emit(ql, q: ops.QParam) -> int:
    """operand form the producer of QuantLayer `ql` writes under quantizer q: 2 = u8 codes (kind::i8 GEMM),
    1 = integer (code - zp) in fp16 (exact kind::f16 GEMM), 0 = de-quantised fp16."""
    if not _exact(q):
        return 0
    return 2 if (USE_I8 and ql is not None and ql.i8_ok(q)) else 1


# implicit-GEMM 3x3 convolution for per-tensor (scalar) activation scales on the kind::i8 path: the producer writes the
# NHWC u8 codes ONCE ([M, C], what a 1x1 conv's producer writes) and the GEMM gathers the 9 taps itself through a 4-D
# TMA map -- no 9x im2col matrix in HBM.  K-wise / row-wise group scales quantise the UNFOLDED view (reference
# quant_layer.py:630-641: up to 9 different codes per input element), so they keep the im2col producer.
IMPLICIT_CONV = os.environ.get("DGQ_IMPLICIT_CONV", "1") != "0"


def _implicit_ok(ql, q: ops.QParam, h: int, w: int, k: int, s: int) -> bool:
    if not (IMPLICIT_CONV and k == 3 and s == 1 and q.mode == ops.Q_SCALAR and not ql.pad_quantized):
        return False
    if _emit(ql, q) != 2 or h < 2 or w < 2:
        return False
    bw = min(w, 16)
    bh = min(h, 128 // bw)
    return (bw & (bw - 1)) == 0 and (bh & (bh - 1)) == 0 and w % bw == 0 and h % bh == 0 and 128 % (bw * bh) == 0


def _gemm(ql, a_op: torch.Tensor, q: ops.QParam = ops.NOQ, *, temb=None, rows_per_batch=0, resid=None,
          want_f32=None, conv_geom=None, **fused):
    """qGEMM of QuantLayer `ql` on the operand its producer wrote under quantizer `q` (u8 codes: kind::i8).
    conv_geom = (b, h, w, c): a_op is the NHWC code tensor of an implicit 3x3 convolution."""
    i8 = a_op.dtype == torch.uint8
    pack = ql.packed(geglu=fused.get("epi") == ops.EPI_GEGLU, i8=i8)
    operand, scale, bias, n_pad = pack[:4]
    if want_f32 is None:
        want_f32 = ops.ACT_DTYPE == torch.float32
    ex = _exact(q) or i8
    if i8:
        fused = dict(fused, colsum=pack[4], b_off=pack[5], row_zp=q.zp)
        if conv_geom is not None:
            fused["conv"] = tuple(conv_geom) + (pack[6],)
    return ops.gemm(a_op, operand, n_pad, scale=scale, bias=bias, temb=temb, rows_per_batch=rows_per_batch,
                    resid=resid, want_f32=want_f32, k=operand.shape[1],
                    row_scale=q.delta if ex else None, row_period=q.period if ex else 1, **fused)


def conv(ql, x: Act, *, x2: Optional[Act] = None, upsample: bool = False, gn=None, act: int = 0,
         temb: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None) -> Act:
    """QuantLayer(nn.Conv2d) on an NHWC activation (optionally the concat of two, upsampled x2,
    GroupNorm+SiLU'd) -- one producer launch + one qGEMM launch."""
    dev = x.t.device
    h, w = (x.h * 2, x.w * 2) if upsample else (x.h, x.w)
    k, s = ql.ksize, ql.stride
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    q = ql.act_qparam(dev)
    if _implicit_ok(ql, q, h, w, k, s):
        a_op = ops.act_producer(x.t, src1=None if x2 is None else x2.t, batch=x.b, h=h, w=w, upsample=upsample,
                                ksize=1, stride=1, gn=gn, act=act, q=q, emit_int=2)
        out = _gemm(ql, a_op, q, temb=temb, rows_per_batch=ho * wo, resid=resid,
                    conv_geom=(x.b, h, w, a_op.shape[1]))
        return Act(out, x.b, ho, wo)
    a_op = ops.act_producer(x.t, src1=None if x2 is None else x2.t, batch=x.b, h=h, w=w, upsample=upsample,
                            ksize=k, stride=s, gn=gn, act=act, q=q, pad_quantized=ql.pad_quantized,
                            emit_int=_emit(ql, q))
    out = _gemm(ql, a_op, q, temb=temb, rows_per_batch=ho * wo, resid=resid)
    return Act(out, x.b, ho, wo)


def linear(ql, a_op: torch.Tensor, q: ops.QParam = ops.NOQ, *, resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qGEMM on an operand produced with emit_int=EXACT_INT under quantizer q (= ql.act_qparam)."""
    return _gemm(ql, a_op, q, resid=resid)


def quant_rows(x: torch.Tensor, ql) -> Tuple[torch.Tensor, ops.QParam]:
    """quantize rows of x for QuantLayer ql -> (operand, q)"""
    q = ql.act_qparam(x.device)
    return ops.row_quant(x, [q], emit_int=_emit(ql, q))[0], q


def quant_layer_forward(ql, x: torch.Tensor) -> torch.Tensor:
    """Stand-alone QuantLayer.forward on torch-layout tensors (reference quant_layer.py:626-661).
    fp32 inputs are quantised from fp32 (bit-exact codes); the GEMM result is returned in fp32."""
    dev = x.device
    q = ql.act_qparam(dev)
    n = ql.out_features
    if ql.is_conv:
        b, c, h, w = x.shape
        src = x.detach().permute(0, 2, 3, 1)
        if src.dtype != torch.float32:
            src = src.float()
        if c % 8:
            src = torch.nn.functional.pad(src, (0, 8 - c % 8))
        src = src.contiguous()
        k, s = ql.ksize, ql.stride
        if q.mode == ops.Q_KWISE and c % 8:
            raise NotImplementedError("K-wise scales need a channel count that is a multiple of 8")
        pad = k // 2
        ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
        if _implicit_ok(ql, q, h, w, k, s) and c % 16 == 0:
            a_op = ops.act_producer(src, batch=b, h=h, w=w, ksize=1, stride=1, q=q, emit_int=2)
            y = _gemm(ql, a_op, q, want_f32=True, conv_geom=(b, h, w, c))
        else:
            a_op = ops.act_producer(src, batch=b, h=h, w=w, ksize=k, stride=s, q=q, pad_quantized=ql.pad_quantized,
                                    emit_int=_emit(ql, q))
            y = _gemm(ql, a_op, q, want_f32=True)
        return y[:, :n].reshape(b, ho, wo, n).permute(0, 3, 1, 2).contiguous().to(x.dtype)
    shp = x.shape
    x2 = x.detach().reshape(-1, shp[-1])
    if x2.dtype not in (torch.float32, torch.float16):
        x2 = x2.float()
    x2 = x2.contiguous()
    if q.mode == ops.Q_ROWWISE and (x.dim() != 3 or q.period != shp[-2]):
        raise ValueError(f"row-wise scales for {q.period} tokens do not fit an input of shape {tuple(shp)}")
    a_op = ops.row_quant(x2, [q], emit_int=_emit(ql, q))[0]
    y = _gemm(ql, a_op, q, want_f32=True)
    return y[:, :n].reshape(*shp[:-1], n).to(x.dtype)


# ------------------------------------------------------------------------------------------
# blocks
# ------------------------------------------------------------------------------------------
def _gn(norm, x: Act, x2: Optional[Act] = None):
    mean, rstd = ops.gn_stats(x.t, None if x2 is None else x2.t, x.b, x.rows, norm.eps)
    return (mean, rstd, _f32(norm.weight), _f32(norm.bias))


_F32_CACHE = {}


def _f32(p: torch.Tensor) -> torch.Tensor:
    """fp32 contiguous view of a (possibly half) norm parameter, cached per version."""
    if p.dtype == torch.float32 and p.is_contiguous():
        return p.detach()
    key = (id(p), p._version)
    hit = _F32_CACHE.get(id(p))
    if hit is None or hit[0] != key:
        hit = (key, p.detach().float().contiguous())
        _F32_CACHE[id(p)] = hit
    return hit[1]


def time_mlp(emb_mod, x: torch.Tensor) -> torch.Tensor:
    """TimestepEmbedding: linear_1 -> SiLU -> linear_2 on [B, C] (2-D inputs: scalar quantizers only)."""
    h = linear(emb_mod.linear_1, *quant_rows(x, emb_mod.linear_1))
    h = ops.silu(h)
    return linear(emb_mod.linear_2, *quant_rows(h, emb_mod.linear_2))


def temb_prefetch(unet, silu_emb: torch.Tensor) -> dict:
    """time_emb_proj(SiLU(temb)) of every resnet on a side stream: ~2 tiny launches per resnet that depend on
    the time embedding only."""
    out = {}
    if not (OVERLAP and OVERLAP_MASK & 4):
        return out
    res = []
    for blk in unet.down_blocks:
        res += list(blk.resnets)
    res += list(unet.mid_block.resnets)
    for blk in unet.up_blocks:
        res += list(blk.resnets)
    with _Fork(silu_emb.device, 3) as f:
        for r in res:
            out[id(r)] = linear(r.time_emb_proj, *quant_rows(silu_emb, r.time_emb_proj))
    out["_fork"] = f
    return out


def resnet(blk, x: Act, silu_emb: torch.Tensor, x2: Optional[Act] = None, te_cache: Optional[dict] = None) -> Act:
    """QuantResnetBlock2D.forward (reference quant_block.py:98-119); x2 = skip tensor to concat."""
    dev = x.t.device
    te = te_cache.get(id(blk)) if te_cache else None
    if te is not None:
        f = te_cache.pop("_fork", None)
        if f is not None:                 # first use: all projections are one side-stream batch
            f.join(*[v for v in te_cache.values()])
    else:
        te = linear(blk.time_emb_proj, *quant_rows(silu_emb, blk.time_emb_proj))
    fsc = None
    if blk.conv_shortcut is not None:
        if OVERLAP and OVERLAP_MASK & 8:  # 1x1 shortcut conv beside GroupNorm statistics + conv1
            with _Fork(dev, 4) as fsc:
                sc = conv(blk.conv_shortcut, x, x2=x2).t
        else:
            sc = conv(blk.conv_shortcut, x, x2=x2).t
    else:
        if x2 is not None:
            raise ValueError("a concatenated input needs conv_shortcut")
        sc = x.t
    h = conv(blk.conv1, x, x2=x2, gn=_gn(blk.norm1, x, x2), act=1, temb=te)
    gn2 = _gn(blk.norm2, h)
    if fsc is not None:
        fsc.join(sc)
    return conv(blk.conv2, h, gn=gn2, act=1, resid=sc)


def _attn_qparam(qt, attn, device):
    """aqtizer_q/k/v parameters: 4-D (B,H,T,D) inputs, so (1,1,X) follows D and (1,X,1) follows T."""
    if not attn.use_aq:
        return ops.NOQ
    return qt.qparam(device)


FUSE_EPILOGUES = True  # GEGLU / head-split+quantize / to_out quantize inside the producing kernels


def _map_args(attn, dev, sp):
    """keyword arguments of ops.attention for attn's softmax-map quantizer (aqtizer_w)."""
    wq = attn.aqtizer_w
    if hasattr(wq, "real_time"):  # T2ILogQuantizer
        return dict(map_mode=ops.MAP_LOG2, real_time=wq.real_time, start_peak=sp,
                    delta=None if wq.real_time else wq.static_delta(dev), qmax=float(wq.level - 1))
    return dict(map_mode=ops.MAP_UNIFORM, start_peak=sp, delta=wq.qparam(dev).delta,  # always_zero uniform
                qmax=float(wq.level - 1))


# Score operands in exact form (reference sd.py:171-183 computes q_hat . k_hat in fp32): Q = the bare integers
# code - zp of aqtizer_q, K = every remaining scale folded in (aqtizer_k's delta in any layout, aqtizer_q's per-channel
# delta) and split into an fp16 (hi | lo) pair, so the tensor-core product carries ~22 bits and the softmax-map codes
# follow the reference's down to its own fp32 rounding (tests/test_layerwise_gpu.py reports the residual rate).
# DGQ_ATTN_SPLIT=0 restores the fp16-rounded operands of round 1 (A/B runs).
ATTN_SPLIT = os.environ.get("DGQ_ATTN_SPLIT", "1") != "0"


def attn_plan(qq: ops.QParam, dp: int) -> dict:
    """how the Q / K operands of an attention are written under Q quantizer `qq`: q_int (bare integers), q_scale /
    q_period (scalar / per-token delta of Q, applied by the kernel through alpha), kfold (per-channel delta of Q,
    folded into K), split (K as an fp16 hi | lo pair)."""
    plan = dict(q_int=False, q_scale=None, q_period=1, kfold=None, split=False)
    if not ATTN_SPLIT or qq.mode == ops.Q_NONE or not qq.int_ok:
        return plan
    plan.update(q_int=True, split=True)
    if qq.mode == ops.Q_KWISE:
        plan["kfold"] = qq.delta
    else:
        plan["q_scale"], plan["q_period"] = qq.delta, qq.period
    return plan


def _attn_plan(attn, dev, use_aq: bool, dp: int) -> dict:
    return attn_plan(_attn_qparam(attn.aqtizer_q, attn, dev) if use_aq else ops.NOQ, dp)


def attention_from_projections(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, b: int, t: int, s: int, heads: int,
                               d: int, qq: ops.QParam, qk: ops.QParam, qv: ops.QParam, *, start_peak: bool = False,
                               **attn_args):
    """Stand-alone form of the attention path (separate dgq_qkv_pack launches instead of the fused GEMM epilogues):
    projections q [b*t, heads*d], k / v [b*s, heads*d] under the quantizers qq / qk / qv -> ops.attention(...)."""
    dp = (d + 63) // 64 * 64
    plan = attn_plan(qq, dp)
    qo = ops.qkv_pack(q, b, t, heads, d, dp, q=qq, emit_int=plan["q_int"])
    ko = ops.qkv_pack(k, b, s, heads, d, dp, skip_first=start_peak, q=qk, kfold=plan["kfold"], split=plan["split"])
    vo = ops.qkv_pack(v, b, s, heads, d, dp, transpose=True, q=qv)
    return ops.attention(qo, ko, vo, d, start_peak=start_peak, q_scale=plan["q_scale"], q_period=plan["q_period"],
                         k_split=plan["split"], **attn_args)


def attention_core(attn, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, b: int, t: int, s: int, *,
                   want_codes: bool = False, out_dtype=None):
    """q [b*t, C], k/v [b*s, C] -> [b*t, C]: quantise + head split, fused two-pass attention
    (stand-alone / unfused form).  want_codes: also the integer codes of the softmax map (verification),
    returned as (out, codes)."""
    dev = q.device
    use_aq = bool(getattr(attn, "use_aq", False))
    sp = bool(getattr(attn, "start_peak", False)) and use_aq
    qs = [(_attn_qparam(getattr(attn, n), attn, dev) if use_aq else ops.NOQ) for n in ("aqtizer_q", "aqtizer_k", "aqtizer_v")]
    margs = _map_args(attn, dev, sp) if use_aq else dict(map_mode=ops.MAP_NONE)
    margs.pop("start_peak", None)
    res = attention_from_projections(q, k, v, b, t, s, attn.num_heads, attn.head_dim, *qs, start_peak=sp,
                                     out_dtype=out_dtype or ops.ACT_DTYPE, want_codes=want_codes, **margs)
    return (res[0], res[2]) if want_codes else res[0]


OVERLAP = True  # independent launches on forked streams: Q/K/V projections of one attention side by side
                # (3 x 4.3 tile waves pack into 13 instead of 15), cross-attention K/V of ALL blocks on a side
                # stream from the start of the call (they depend on the prompt embedding only)
OVERLAP_MASK = int(os.environ.get("DGQ_OVERLAP", "15"))   # 1 q/k/v, 2 cross K/V, 4 time-embedding, 8 shortcut (A/B runs)
_SIDE: dict = {}


def _side_streams(dev, n: int):
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    pool = _SIDE.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


class _Fork:
    """`with _Fork(dev, i) as f:` issues the body on side stream i, ordered after everything already on the
    calling stream; f.join(*tensors) orders the calling stream after the body and hands tensors that were
    allocated inside it over to the calling stream (caching-allocator bookkeeping)."""

    def __init__(self, dev, idx: int):
        self.main = torch.cuda.current_stream()
        self.side = _side_streams(dev, idx + 1)[idx]
        self.done = None

    def __enter__(self):
        ev = torch.cuda.Event()
        ev.record(self.main)
        self.side.wait_event(ev)
        self._ctx = torch.cuda.stream(self.side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self.done = torch.cuda.Event()
        self.done.record(self.side)
        return self._ctx.__exit__(*exc)

    def join(self, *tensors):
        torch.cuda.current_stream().wait_event(self.done)
        for t in tensors:
            t.record_stream(torch.cuda.current_stream())


def _qkv_gemm(attn, which: int, x, qin, q2, b, tok, dst, sp, plan):
    heads, d = attn.num_heads, attn.head_dim
    dp = (d + 63) // 64 * 64
    ql = (attn.to_q, attn.to_k, attn.to_v)[which]
    extra = {}
    if which == 0:
        extra = dict(q2_emit_int=int(plan["q_int"]))
    elif which == 1:
        extra = dict(kfold=plan["kfold"], k_split=plan["split"])
    _gemm(ql, x, qin, epi=ops.EPI_QKV, q2=q2, out=dst,
          qkv=(heads, d, dp, tok, (tok + 7) // 8 * 8, which == 2, sp and which == 1), **extra)


def attention(attn, xq, xk, xv, qs, b: int, t: int, s: int, resid: Optional[torch.Tensor], kv=None) -> torch.Tensor:
    """Attention_forward given the three already-quantised projection inputs (operands xq/xk/xv
    written under quantizers qs = [q_to_q, q_to_k, q_to_v]).  kv = (K, V^T, event): projections already
    done on a side stream (cross_kv_prefetch)."""
    if not FUSE_EPILOGUES:
        q = linear(attn.to_q, xq, qs[0])
        k = linear(attn.to_k, xk, qs[1])
        v = linear(attn.to_v, xv, qs[2])
        o = attention_core(attn, q, k, v, b, t, s)
        return linear(attn.to_out[0], *quant_rows(o, attn.to_out[0]), resid=resid)
    dev = xq.device
    heads, d = attn.num_heads, attn.head_dim
    dp = (d + 63) // 64 * 64
    use_aq = bool(getattr(attn, "use_aq", False))
    sp = bool(getattr(attn, "start_peak", False)) and use_aq
    aq = [(_attn_qparam(getattr(attn, n), attn, dev) if use_aq else ops.NOQ) for n in ("aqtizer_q", "aqtizer_k", "aqtizer_v")]
    main = torch.cuda.current_stream()
    plan = _attn_plan(attn, dev, use_aq, dp)
    dq = ops.qkv_dest(b, t, heads, d, dp, False, dev)
    if kv is not None:
        dk, dv, ev = kv
        _qkv_gemm(attn, 0, xq, qs[0], aq[0], b, t, dq, sp, plan)
        main.wait_event(ev)
    else:
        dk = ops.qkv_dest(b, s, heads, d, dp, False, dev, split=plan["split"])
        dv = ops.qkv_dest(b, s, heads, d, dp, True, dev)
        if OVERLAP and OVERLAP_MASK & 1:
            s1, s2 = _side_streams(dev, 2)
            fork = torch.cuda.Event()
            fork.record(main)
            for st, which, x, dst in ((s1, 1, xk, dk), (s2, 2, xv, dv)):
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    _qkv_gemm(attn, which, x, qs[which], aq[which], b, s, dst, sp, plan)
            _qkv_gemm(attn, 0, xq, qs[0], aq[0], b, t, dq, sp, plan)
            for st in (s1, s2):
                ev = torch.cuda.Event()
                ev.record(st)
                main.wait_event(ev)
        else:
            _qkv_gemm(attn, 0, xq, qs[0], aq[0], b, t, dq, sp, plan)
            _qkv_gemm(attn, 1, xk, qs[1], aq[1], b, s, dk, sp, plan)
            _qkv_gemm(attn, 2, xv, qs[2], aq[2], b, s, dv, sp, plan)
    qo = attn.to_out[0].act_qparam(dev)
    margs = _map_args(attn, dev, sp) if use_aq else dict(map_mode=ops.MAP_NONE)
    o, _ = ops.attention(dq, dk, dv, d, out_q=qo, out_emit_int=_emit(attn.to_out[0], qo), q_scale=plan["q_scale"],
                         q_period=plan["q_period"], k_split=plan["split"], **margs)
    return linear(attn.to_out[0], o, qo, resid=resid)


def cross_kv_prefetch(unet, ctx: torch.Tensor) -> dict:
    """K / V^T of every cross-attention (attn2.to_k / to_v on the prompt embedding, each under its own
    activation quantizers) on a side stream, forked at the start of the UNet call.  These ~210 small
    launches (77-token GEMMs: 25-50 tiles) are latency-bound; off the critical path they fill the tile-wave
    tails of the main stream's kernels.  Destination buffers are allocated on the calling stream."""
    out = {}
    if not (OVERLAP and OVERLAP_MASK & 2 and FUSE_EPILOGUES) or ctx is None:
        return out
    t2d = []                              # Transformer2DModels in execution order
    for blk in unet.down_blocks:
        t2d += list(getattr(blk, "attentions", None) or [])
    t2d += list(unet.mid_block.attentions)
    for blk in unet.up_blocks:
        t2d += list(getattr(blk, "attentions", None) or [])
    blocks = [tb for m in t2d for tb in m.transformer_blocks]
    if not blocks:
        return out
    dev = ctx.device
    main = torch.cuda.current_stream()
    cx, cb, s = _ctx_operand(ctx)
    side = _side_streams(dev, 3)[2]
    dests, plans = [], []
    for blk in blocks:
        a2 = blk.attn2
        heads, d = a2.num_heads, a2.head_dim
        dp = (d + 63) // 64 * 64
        plans.append(_attn_plan(a2, dev, bool(getattr(a2, "use_aq", False)), dp))
        dests.append((ops.qkv_dest(cb, s, heads, d, dp, False, dev, split=plans[-1]["split"]),
                      ops.qkv_dest(cb, s, heads, d, dp, True, dev)))
    fork = torch.cuda.Event()
    fork.record(main)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        for blk, (dk, dv), plan in zip(blocks, dests, plans):
            a2 = blk.attn2
            use_aq = bool(getattr(a2, "use_aq", False))
            sp = bool(getattr(a2, "start_peak", False)) and use_aq
            qs = [None, a2.to_k.act_qparam(dev), a2.to_v.act_qparam(dev)]
            aq = [None] + [(_attn_qparam(getattr(a2, n), a2, dev) if use_aq else ops.NOQ) for n in ("aqtizer_k", "aqtizer_v")]
            xkv = ops.row_quant(cx, qs[1:], emit_int=[_emit(a2.to_k, qs[1]), _emit(a2.to_v, qs[2])])
            _qkv_gemm(a2, 1, xkv[0], qs[1], aq[1], cb, s, dk, sp, plan)
            _qkv_gemm(a2, 2, xkv[1], qs[2], aq[2], cb, s, dv, sp, plan)
            ev = torch.cuda.Event()
            ev.record(side)
            out[id(a2)] = (dk, dv, ev)
    if cx.data_ptr() != ctx.data_ptr():
        cx.record_stream(side)            # a converted copy (bf16 / strided ctx) allocated on the calling stream
    out["_ctx_operand"] = cx              # ... and kept alive until the last attn2 has joined
    return out


def _ctx_operand(ctx: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    b, s, c = ctx.shape
    x = ctx.detach().reshape(b * s, c)
    if x.dtype not in (torch.float32, torch.float16):
        x = x.float()
    return x.contiguous(), b, s


def transformer_block(blk, h: Act, ctx: Optional[torch.Tensor], kv_cache: Optional[dict] = None) -> Act:
    """QuantBasicTransformerBlock.forward (reference quant_block.py:165-186)."""
    dev = h.t.device
    b, t = h.b, h.rows
    a1, a2, ff = blk.attn1, blk.attn2, blk.ff
    qs = [a1.to_q.act_qparam(dev), a1.to_k.act_qparam(dev), a1.to_v.act_qparam(dev)]
    xs = ops.ln_quant(h.t, _f32(blk.norm1.weight), _f32(blk.norm1.bias), blk.norm1.eps, qs,
                      emit_int=[_emit(l, q) for l, q in zip((a1.to_q, a1.to_k, a1.to_v), qs)])
    x = attention(a1, xs[0], xs[1], xs[2], qs, b, t, t, resid=h.t)
    if ctx is not None:
        cx, cb, s = _ctx_operand(ctx)
        qs = [a2.to_q.act_qparam(dev), a2.to_k.act_qparam(dev), a2.to_v.act_qparam(dev)]
        xq = ops.ln_quant(x, _f32(blk.norm2.weight), _f32(blk.norm2.bias), blk.norm2.eps, qs[:1],
                          emit_int=[_emit(a2.to_q, qs[0])])[0]
        kv = kv_cache.get(id(a2)) if kv_cache else None
        if kv is not None:
            x = attention(a2, xq, None, None, qs, b, t, s, resid=x, kv=kv)
        else:
            xkv = ops.row_quant(cx, qs[1:], emit_int=[_emit(a2.to_k, qs[1]), _emit(a2.to_v, qs[2])])
            x = attention(a2, xq, xkv[0], xkv[1], qs, b, t, s, resid=x)
    else:
        qs = [a2.to_q.act_qparam(dev), a2.to_k.act_qparam(dev), a2.to_v.act_qparam(dev)]
        xs = ops.ln_quant(x, _f32(blk.norm2.weight), _f32(blk.norm2.bias), blk.norm2.eps, qs,
                          emit_int=[_emit(l, q) for l, q in zip((a2.to_q, a2.to_k, a2.to_v), qs)])
        x = attention(a2, xs[0], xs[1], xs[2], qs, b, t, t, resid=x)
    proj, out = ff.net[0].proj, ff.net[2]
    qp = proj.act_qparam(dev)
    x3 = ops.ln_quant(x, _f32(blk.norm3.weight), _f32(blk.norm3.bias), blk.norm3.eps, [qp],
                      emit_int=[_emit(proj, qp)])[0]
    qo = out.act_qparam(dev)
    if FUSE_EPILOGUES and proj.out_features % 64 == 0:
        g = _gemm(proj, x3, qp, epi=ops.EPI_GEGLU, q2=qo, q2_emit_int=_emit(out, qo))
    else:
        g = ops.geglu_quant(linear(proj, x3, qp), qo, emit_int=_emit(out, qo))
    x = linear(out, g, qo, resid=x)
    return Act(x, h.b, h.h, h.w)


def transformer2d(mod, x: Act, ctx: Optional[torch.Tensor], kv_cache: Optional[dict] = None) -> Act:
    """Transformer2DModel.forward (sd.py:283-305 conv proj; sdxl.py:306-326 linear proj): in NHWC
    both are the same GEMM, and the NCHW<->token permutes of the reference disappear."""
    dev = x.t.device
    if mod.proj_in.is_conv:
        h = conv(mod.proj_in, x, gn=_gn(mod.norm, x))
    else:
        qi = mod.proj_in.act_qparam(dev)
        a_op = ops.act_producer(x.t, batch=x.b, h=x.h, w=x.w, ksize=1, gn=_gn(mod.norm, x), q=qi,
                                emit_int=_emit(mod.proj_in, qi))
        h = Act(linear(mod.proj_in, a_op, qi), x.b, x.h, x.w)
    for blk in mod.transformer_blocks:
        h = transformer_block(blk, h, ctx, kv_cache)
    if mod.proj_out.is_conv:
        return conv(mod.proj_out, h, resid=x.t)
    y = linear(mod.proj_out, *quant_rows(h.t, mod.proj_out), resid=x.t)
    return Act(y, x.b, x.h, x.w)


# ------------------------------------------------------------------------------------------
# UNet
# ------------------------------------------------------------------------------------------
TAPS: Optional[list] = None  # debugging: set to [] to record (name, NCHW fp32) after every block


def _tap(name: str, a) -> None:
    if TAPS is not None:
        TAPS.append((name, a.float().clone() if torch.is_tensor(a) else act_to_nchw(a)))


def unet_forward(unet, sample: torch.Tensor, timesteps: torch.Tensor, ctx: torch.Tensor,
                 added: Optional[dict] = None) -> torch.Tensor:
    """UNet2DConditionModel.forward (sd.py:546-620, sdxl.py:558-631) -> NCHW tensor like `sample`."""
    dev = sample.device
    bsz = sample.shape[0]
    t = timesteps.reshape(-1).to(device=dev, dtype=torch.float32).expand(bsz).contiguous()
    emb = time_mlp(unet.time_embedding, ops.timestep_embedding(t, unet.time_proj.num_channels, f32=True))
    if hasattr(unet, "add_embedding"):
        te = ops.timestep_embedding(added["time_ids"].to(dev).flatten(), unet.add_time_proj.num_channels, f32=True)
        add = torch.cat([added["text_embeds"].to(dev).float(), te.reshape(bsz, -1)], dim=-1).contiguous()
        emb = ops.add(emb, time_mlp(unet.add_embedding, add))
    silu_emb = ops.silu(emb)  # nonlinearity(temb) is the same tensor for every resnet

    _tap("emb", emb)
    kvc = cross_kv_prefetch(unet, ctx)
    tec = temb_prefetch(unet, silu_emb)
    h = conv(unet.conv_in, act_from_nchw(sample))
    _tap("conv_in", h)
    skips: List[Act] = [h]
    for i, blk in enumerate(unet.down_blocks):
        attns = getattr(blk, "attentions", None)
        for j, res in enumerate(blk.resnets):
            h = resnet(res, h, silu_emb, te_cache=tec)
            _tap(f"down{i}.res{j}", h)
            if attns is not None:
                h = transformer2d(attns[j], h, ctx, kvc)
                _tap(f"down{i}.attn{j}", h)
            skips.append(h)
        if getattr(blk, "downsamplers", None) is not None:
            h = conv(blk.downsamplers[0].conv, h)
            skips.append(h)

    mid = unet.mid_block
    h = resnet(mid.resnets[0], h, silu_emb, te_cache=tec)
    for attn, res in zip(mid.attentions, mid.resnets[1:]):
        h = transformer2d(attn, h, ctx, kvc)
        h = resnet(res, h, silu_emb, te_cache=tec)
    _tap("mid", h)

    for i, blk in enumerate(unet.up_blocks):
        attns = getattr(blk, "attentions", None)
        for j, res in enumerate(blk.resnets):
            h = resnet(res, h, silu_emb, x2=skips.pop(), te_cache=tec)  # torch.cat([h, skip], 1) fused into the producers
            _tap(f"up{i}.res{j}", h)
            if attns is not None:
                h = transformer2d(attns[j], h, ctx, kvc)
                _tap(f"up{i}.attn{j}", h)
        if getattr(blk, "upsamplers", None) is not None:
            for up in blk.upsamplers:
                h = conv(up.conv, h, upsample=True)

    out = conv(unet.conv_out, h, gn=_gn(unet.conv_norm_out, h), act=1)
    return act_to_nchw(out, c=unet.conv_out.out_features, dtype=sample.dtype)
