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


def _exact(q: ops.QParam) -> bool:
    return EXACT_INT and q.exact


def _gemm(ql, a_op: torch.Tensor, q: ops.QParam = ops.NOQ, *, temb=None, rows_per_batch=0, resid=None,
          want_f32=None, **fused):
    """qGEMM of QuantLayer `ql` on the operand its producer wrote under quantizer `q`."""
    operand, scale, bias, n_pad = ql.packed(geglu=fused.get("epi") == ops.EPI_GEGLU)
    if want_f32 is None:
        want_f32 = ops.ACT_DTYPE == torch.float32
    ex = _exact(q)
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
    a_op = ops.act_producer(x.t, src1=None if x2 is None else x2.t, batch=x.b, h=h, w=w, upsample=upsample,
                            ksize=k, stride=s, gn=gn, act=act, q=q, pad_quantized=ql.pad_quantized,
                            emit_int=_exact(q))
    out = _gemm(ql, a_op, q, temb=temb, rows_per_batch=ho * wo, resid=resid)
    return Act(out, x.b, ho, wo)


def linear(ql, a_op: torch.Tensor, q: ops.QParam = ops.NOQ, *, resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qGEMM on an operand produced with emit_int=EXACT_INT under quantizer q (= ql.act_qparam)."""
    return _gemm(ql, a_op, q, resid=resid)


def quant_rows(x: torch.Tensor, ql) -> Tuple[torch.Tensor, ops.QParam]:
    """quantize rows of x for QuantLayer ql -> (operand, q)"""
    q = ql.act_qparam(x.device)
    return ops.row_quant(x, [q], emit_int=EXACT_INT)[0], q


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
        a_op = ops.act_producer(src, batch=b, h=h, w=w, ksize=k, stride=s, q=q, pad_quantized=ql.pad_quantized,
                                emit_int=_exact(q))
        pad = k // 2
        ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
        y = _gemm(ql, a_op, q, want_f32=True)
        return y[:, :n].reshape(b, ho, wo, n).permute(0, 3, 1, 2).contiguous().to(x.dtype)
    shp = x.shape
    x2 = x.detach().reshape(-1, shp[-1])
    if x2.dtype not in (torch.float32, torch.float16):
        x2 = x2.float()
    x2 = x2.contiguous()
    if q.mode == ops.Q_ROWWISE and (x.dim() != 3 or q.period != shp[-2]):
        raise ValueError(f"row-wise scales for {q.period} tokens do not fit an input of shape {tuple(shp)}")
    a_op = ops.row_quant(x2, [q], emit_int=EXACT_INT)[0]
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


def resnet(blk, x: Act, silu_emb: torch.Tensor, x2: Optional[Act] = None) -> Act:
    """QuantResnetBlock2D.forward (reference quant_block.py:98-119); x2 = skip tensor to concat."""
    dev = x.t.device
    te = linear(blk.time_emb_proj, *quant_rows(silu_emb, blk.time_emb_proj))
    h = conv(blk.conv1, x, x2=x2, gn=_gn(blk.norm1, x, x2), act=1, temb=te)
    if blk.conv_shortcut is not None:
        sc = conv(blk.conv_shortcut, x, x2=x2).t
    else:
        if x2 is not None:
            raise ValueError("a concatenated input needs conv_shortcut")
        sc = x.t
    return conv(blk.conv2, h, gn=_gn(blk.norm2, h), act=1, resid=sc)


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


def attention_core(attn, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, b: int, t: int, s: int) -> torch.Tensor:
    """q [b*t, C], k/v [b*s, C] -> [b*t, C]: quantise + head split, fused two-pass attention
    (stand-alone / unfused form: separate dgq_qkv_pack launches)."""
    dev = q.device
    heads, d = attn.num_heads, attn.head_dim
    dp = (d + 63) // 64 * 64
    use_aq = bool(getattr(attn, "use_aq", False))
    sp = bool(getattr(attn, "start_peak", False)) and use_aq
    qq = ops.qkv_pack(q, b, t, heads, d, dp, q=_attn_qparam(attn.aqtizer_q, attn, dev) if use_aq else ops.NOQ)
    kk = ops.qkv_pack(k, b, s, heads, d, dp, skip_first=sp,
                      q=_attn_qparam(attn.aqtizer_k, attn, dev) if use_aq else ops.NOQ)
    vv = ops.qkv_pack(v, b, s, heads, d, dp, transpose=True,
                      q=_attn_qparam(attn.aqtizer_v, attn, dev) if use_aq else ops.NOQ)
    if not use_aq:
        out, _ = ops.attention(qq, kk, vv, d, map_mode=ops.MAP_NONE, out_dtype=ops.ACT_DTYPE)
        return out
    out, _ = ops.attention(qq, kk, vv, d, out_dtype=ops.ACT_DTYPE, **_map_args(attn, dev, sp))
    return out


def attention(attn, xq, xk, xv, qs, b: int, t: int, s: int, resid: Optional[torch.Tensor]) -> torch.Tensor:
    """Attention_forward given the three already-quantised projection inputs (operands xq/xk/xv
    written under quantizers qs = [q_to_q, q_to_k, q_to_v])."""
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
    ops_qkv = []
    for ql, x, qin, q2, tok, tr, skip in ((attn.to_q, xq, qs[0], aq[0], t, False, False),
                                         (attn.to_k, xk, qs[1], aq[1], s, False, sp),
                                         (attn.to_v, xv, qs[2], aq[2], s, True, False)):
        dst = ops.qkv_dest(b, tok, heads, d, dp, tr, dev)
        _gemm(ql, x, qin, epi=ops.EPI_QKV, q2=q2, out=dst,
              qkv=(heads, d, dp, tok, (tok + 7) // 8 * 8, tr, skip))
        ops_qkv.append(dst)
    qo = attn.to_out[0].act_qparam(dev)
    margs = _map_args(attn, dev, sp) if use_aq else dict(map_mode=ops.MAP_NONE)
    o, _ = ops.attention(ops_qkv[0], ops_qkv[1], ops_qkv[2], d, out_q=qo, out_emit_int=_exact(qo), **margs)
    return linear(attn.to_out[0], o, qo, resid=resid)


def _ctx_operand(ctx: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    b, s, c = ctx.shape
    x = ctx.detach().reshape(b * s, c)
    if x.dtype not in (torch.float32, torch.float16):
        x = x.float()
    return x.contiguous(), b, s


def transformer_block(blk, h: Act, ctx: Optional[torch.Tensor]) -> Act:
    """QuantBasicTransformerBlock.forward (reference quant_block.py:165-186)."""
    dev = h.t.device
    b, t = h.b, h.rows
    a1, a2, ff = blk.attn1, blk.attn2, blk.ff
    qs = [a1.to_q.act_qparam(dev), a1.to_k.act_qparam(dev), a1.to_v.act_qparam(dev)]
    xs = ops.ln_quant(h.t, _f32(blk.norm1.weight), _f32(blk.norm1.bias), blk.norm1.eps, qs, emit_int=EXACT_INT)
    x = attention(a1, xs[0], xs[1], xs[2], qs, b, t, t, resid=h.t)
    if ctx is not None:
        cx, cb, s = _ctx_operand(ctx)
        qs = [a2.to_q.act_qparam(dev), a2.to_k.act_qparam(dev), a2.to_v.act_qparam(dev)]
        xq = ops.ln_quant(x, _f32(blk.norm2.weight), _f32(blk.norm2.bias), blk.norm2.eps, qs[:1],
                          emit_int=EXACT_INT)[0]
        xkv = ops.row_quant(cx, qs[1:], emit_int=EXACT_INT)
        x = attention(a2, xq, xkv[0], xkv[1], qs, b, t, s, resid=x)
    else:
        qs = [a2.to_q.act_qparam(dev), a2.to_k.act_qparam(dev), a2.to_v.act_qparam(dev)]
        xs = ops.ln_quant(x, _f32(blk.norm2.weight), _f32(blk.norm2.bias), blk.norm2.eps, qs, emit_int=EXACT_INT)
        x = attention(a2, xs[0], xs[1], xs[2], qs, b, t, t, resid=x)
    proj, out = ff.net[0].proj, ff.net[2]
    qp = proj.act_qparam(dev)
    x3 = ops.ln_quant(x, _f32(blk.norm3.weight), _f32(blk.norm3.bias), blk.norm3.eps, [qp], emit_int=EXACT_INT)[0]
    qo = out.act_qparam(dev)
    if FUSE_EPILOGUES and proj.out_features % 64 == 0:
        g = _gemm(proj, x3, qp, epi=ops.EPI_GEGLU, q2=qo, q2_emit_int=EXACT_INT)
    else:
        g = ops.geglu_quant(linear(proj, x3, qp), qo, emit_int=EXACT_INT)
    x = linear(out, g, qo, resid=x)
    return Act(x, h.b, h.h, h.w)


def transformer2d(mod, x: Act, ctx: Optional[torch.Tensor]) -> Act:
    """Transformer2DModel.forward (sd.py:283-305 conv proj; sdxl.py:306-326 linear proj): in NHWC
    both are the same GEMM, and the NCHW<->token permutes of the reference disappear."""
    dev = x.t.device
    if mod.proj_in.is_conv:
        h = conv(mod.proj_in, x, gn=_gn(mod.norm, x))
    else:
        qi = mod.proj_in.act_qparam(dev)
        a_op = ops.act_producer(x.t, batch=x.b, h=x.h, w=x.w, ksize=1, gn=_gn(mod.norm, x), q=qi,
                                emit_int=_exact(qi))
        h = Act(linear(mod.proj_in, a_op, qi), x.b, x.h, x.w)
    for blk in mod.transformer_blocks:
        h = transformer_block(blk, h, ctx)
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
    h = conv(unet.conv_in, act_from_nchw(sample))
    _tap("conv_in", h)
    skips: List[Act] = [h]
    for i, blk in enumerate(unet.down_blocks):
        attns = getattr(blk, "attentions", None)
        for j, res in enumerate(blk.resnets):
            h = resnet(res, h, silu_emb)
            _tap(f"down{i}.res{j}", h)
            if attns is not None:
                h = transformer2d(attns[j], h, ctx)
                _tap(f"down{i}.attn{j}", h)
            skips.append(h)
        if getattr(blk, "downsamplers", None) is not None:
            h = conv(blk.downsamplers[0].conv, h)
            skips.append(h)

    mid = unet.mid_block
    h = resnet(mid.resnets[0], h, silu_emb)
    for attn, res in zip(mid.attentions, mid.resnets[1:]):
        h = transformer2d(attn, h, ctx)
        h = resnet(res, h, silu_emb)
    _tap("mid", h)

    for i, blk in enumerate(unet.up_blocks):
        attns = getattr(blk, "attentions", None)
        for j, res in enumerate(blk.resnets):
            h = resnet(res, h, silu_emb, x2=skips.pop())  # torch.cat([h, skip], 1) fused into the producers
            _tap(f"up{i}.res{j}", h)
            if attns is not None:
                h = transformer2d(attns[j], h, ctx)
                _tap(f"up{i}.attn{j}", h)
        if getattr(blk, "upsamplers", None) is not None:
            for up in blk.upsamplers:
                h = conv(up.conv, h, upsample=True)

    out = conv(unet.conv_out, h, gn=_gn(unet.conv_norm_out, h), act=1)
    return act_to_nchw(out, c=unet.conv_out.out_features, dtype=sample.dtype)
