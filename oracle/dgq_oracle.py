"""CPU oracle for DGQ's quantized UNet forward path -- TEST INFRASTRUCTURE ONLY.

This module is a plain torch-CPU fp32 restatement of the reference's fake-quant
algorithm.  It is the checker for the CUDA path; nothing under ``dgq_b200/``
imports it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline
legs of ``bench.py`` may import it.

Parity pinning: the reference (ugonfor/DGQ) ships no golden vectors for this
path.  The oracle is pinned against outputs of the
reference itself, executed in the build container by
``tests/golden/make_golden.py`` and committed as fixtures under
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks every function
here against them.

All ``file:line`` citations are relative to the reference tree.

State layout: the oracle is functional.  ``sd`` is a flat dict with exactly the
key names of the reference's ``QuantModel.state_dict()`` (``model.<path>.w``,
``.b``, ``.wqtizer.delta``, ``.wqtizer.zero_point``, optional
``.wqtizer.alpha``, GroupNorm/LayerNorm ``.weight/.bias``); ``act`` is one
``act_k`` dict of the activation checkpoint (``model.<path>.aqtizer.delta`` ...).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# quantizers
# --------------------------------------------------------------------------- #
def uaq_codes(x: Tensor, delta: Tensor, zp: Tensor, level: int) -> Tensor:
    """Integer codes of UniformAffineQuantizer (quant/quant_layer.py:295-297).

    asymmetric => NB=0, PB=level-1.  ``ste_round`` (``:212-213``) equals
    ``round`` in value; torch.round is round-half-to-even; the division is a
    true IEEE division because delta is a tensor.
    """
    return torch.clamp(torch.round(x / delta) + zp, 0, level - 1)


def uaq_fake_quant(x: Tensor, delta: Tensor, zp: Tensor, level: int) -> Tensor:
    """quant/quant_layer.py:297-299."""
    return delta * (uaq_codes(x, delta, zp, level) - zp)


def minmax_scale(x: Tensor, level: int, always_zero: bool = False) -> Tuple[Tensor, Tensor]:
    """Scaler.MINMAX, asymmetric (quant/quant_layer.py:22-38)."""
    x_min, x_max = min(x.min().item(), 0), max(x.max().item(), 0)
    delta = torch.tensor(float(x_max - x_min) / (level - 1))
    if always_zero:
        delta = torch.tensor(float(x_max) / (level - 1))
    if delta < 1e-8:
        delta = torch.tensor(1e-8)
    if always_zero:
        zp = torch.tensor(0.0)
    else:
        zp = torch.round(-torch.tensor(x_min) / delta)
    return delta.to(x.dtype), zp.to(x.dtype)


def channel_minmax_scale(w: Tensor, level: int) -> Tuple[Tensor, Tensor]:
    """Per-out-channel MINMAX init of the weight quantizer
    (quant/quant_layer.py:253-264): one (delta, zp) per w[c]."""
    n = w.shape[0]
    delta = torch.empty(n, dtype=w.dtype)
    zp = torch.empty(n, dtype=w.dtype)
    for c in range(n):
        delta[c], zp[c] = minmax_scale(w[c], level)
    shape = (-1,) + (1,) * (w.dim() - 1)
    return delta.view(shape), zp.view(shape)


def adaround_codes(w: Tensor, delta: Tensor, zp: Tensor, alpha: Tensor, level: int) -> Tensor:
    """AdaRoundQuantizer hard rounding (quant/adaptive_rounding.py:51-70,
    soft_tgt=False :62-63)."""
    x_int = torch.floor(w / delta) + (alpha >= 0).to(w.dtype)
    return torch.clamp(x_int + zp, 0, level - 1)


def weight_codes(sd: Dict[str, Tensor], name: str, level: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(codes, delta, zp) of layer ``name``'s weight; AdaRound when an alpha key
    is present (quant/calibration.py:227-230)."""
    w = sd[name + ".w"]
    delta = sd[name + ".wqtizer.delta"]
    zp = sd[name + ".wqtizer.zero_point"]
    if name + ".wqtizer.alpha" in sd:
        codes = adaround_codes(w, delta, zp, sd[name + ".wqtizer.alpha"], level)
    else:
        codes = uaq_codes(w, delta, zp, level)
    return codes, delta, zp


def t2i_log_codes(x: Tensor, delta: Tensor, level: int) -> Tensor:
    """T2ILogQuantizer codes (quant/quant_layer_text.py:101-103)."""
    return torch.clamp(torch.round(-1 * torch.log2(x / delta)), 0, level - 1)


def t2i_log_fake_quant(x: Tensor, delta: Optional[Tensor], level: int, real_time: bool) -> Tensor:
    """quant/quant_layer_text.py:96-105.  real_time => delta = x.max()."""
    d = x.max() if real_time else delta
    return 2 ** (-1 * t2i_log_codes(x, d, level)) * d


# --------------------------------------------------------------------------- #
# configuration
# --------------------------------------------------------------------------- #
@dataclass
class QConfig:
    """Mirrors the three param dicts of src/inference_qmodel.py:73-87."""
    wbits: int = 8
    abits: int = 8
    use_wq: bool = True
    use_aq: bool = True
    softmax_bits: int = 8
    t2i_log_quant: bool = False
    t2i_real_time: bool = False
    t2i_start_peak: bool = False
    # names of conv layers whose QuantLayer.use_group_num is set (sticky flag,
    # quant/calibration.py:271-278): they run the unfold + matmul path
    group_convs: set = field(default_factory=set)


# --------------------------------------------------------------------------- #
# QuantLayer
# --------------------------------------------------------------------------- #
def quant_layer(x: Tensor, sd: Dict[str, Tensor], act: Optional[Dict[str, Tensor]],
                name: str, cfg: QConfig, *, stride: int = 1, padding: int = 0,
                fp_layer: bool = False) -> Tensor:
    """QuantLayer.forward (quant/quant_layer.py:626-661).

    ``fp_layer`` = conv_in/conv_out after disable_out_quantization
    (quant/quant_model.py:118-124): original weights, no activation quant.
    """
    w = sd[name + ".w"]
    b = sd.get(name + ".b")
    is_conv = w.dim() == 4
    grouped = is_conv and name in cfg.group_convs
    in_shape = x.shape
    if grouped:  # :630-638
        x = F.unfold(x, kernel_size=(w.shape[2], w.shape[3]), dilation=1,
                     padding=padding, stride=stride)
    if cfg.use_aq and not fp_layer and act is not None and (name + ".aqtizer.delta") in act:  # :640-641
        x = uaq_fake_quant(x, act[name + ".aqtizer.delta"], act[name + ".aqtizer.zero_point"],
                           2 ** cfg.abits)
    if cfg.use_wq and not fp_layer:  # :642-644
        codes, delta, zp = weight_codes(sd, name, 2 ** cfg.wbits)
        w = delta * (codes - zp)
    if grouped:  # :649-657 + input_unfolded_pseudo_conv2d :526-574
        co = w.shape[0]
        out = w.view(co, -1) @ x
        ho = (in_shape[2] + 2 * padding - (w.shape[2] - 1) - 1) // stride + 1
        wo = (in_shape[3] + 2 * padding - (w.shape[3] - 1) - 1) // stride + 1
        out = out.view(in_shape[0], co, ho, wo)
        if b is not None:
            out = out + b.view(1, co, 1, 1)
        return out
    if is_conv:
        return F.conv2d(x, w, b, stride=stride, padding=padding)
    return F.linear(x, w, b)


# --------------------------------------------------------------------------- #
# attention
# --------------------------------------------------------------------------- #
def _aq(act, key, x, level):
    if act is None or (key + ".delta") not in act:
        return x
    return uaq_fake_quant(x, act[key + ".delta"], act[key + ".zero_point"], level)


def _map_quant(m: Tensor, act, name: str, cfg: "QConfig") -> Tensor:
    """aqtizer_w on the softmax map: T2ILogQuantizer or always_zero uniform
    (quant/quant_block.py:149-156)."""
    lv = 2 ** cfg.softmax_bits
    if cfg.t2i_log_quant:
        dl = None if cfg.t2i_real_time else act[name + ".aqtizer_w.delta"]
        return t2i_log_fake_quant(m, dl, lv, cfg.t2i_real_time)
    return uaq_fake_quant(m, act[name + ".aqtizer_w.delta"],
                          act[name + ".aqtizer_w.zero_point"], lv)


def attention_core(q: Tensor, k: Tensor, v: Tensor, act, name: str, cfg: QConfig, *, is_cross: bool,
                   return_probs: bool = False) -> Tensor:
    """The part of Attention_forward between the projections (sd.py:171-201): q, k, v are
    (B, H, T, D); returns (B, T, H*D)."""
    b, heads, t, d = q.shape
    start_peak = cfg.t2i_start_peak and is_cross  # quant_block.py:157-158
    level = 2 ** cfg.abits
    if cfg.use_aq:
        q = _aq(act, name + ".aqtizer_q", q, level)
        if start_peak:  # sd.py:176-180
            k = torch.cat([k[..., 0:1, :], _aq(act, name + ".aqtizer_k", k[..., 1:, :], level)], dim=-2)
        else:
            k = _aq(act, name + ".aqtizer_k", k, level)
    scores = torch.matmul(q, k.transpose(-2, -1)) * (d ** -0.5)
    p = torch.softmax(scores, dim=-1)
    if cfg.use_aq:  # sd.py:187-199
        p = p.float()
        if start_peak:
            p = torch.cat([p[..., 0:1], _map_quant(p[..., 1:], act, name, cfg)], dim=-1)
        else:
            p = _map_quant(p, act, name, cfg)
        v = _aq(act, name + ".aqtizer_v", v, level)
    if return_probs:
        return p
    return torch.matmul(p, v).transpose(1, 2).contiguous().view(b, t, heads * d)


def attention(x: Tensor, ctx: Optional[Tensor], sd, act, name: str, cfg: QConfig, *,
              heads: int, is_cross: bool) -> Tensor:
    """Attention.Attention_forward (diffusers_rewrite/sd.py:151-207, sdxl.py:174-229)."""
    src = ctx if ctx is not None else x
    q = quant_layer(x, sd, act, name + ".to_q", cfg)
    k = quant_layer(src, sd, act, name + ".to_k", cfg)
    v = quant_layer(src, sd, act, name + ".to_v", cfg)
    b, t, c = q.shape
    d = c // heads
    q = q.view(b, -1, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    o = attention_core(q, k, v, act, name, cfg, is_cross=is_cross)
    return quant_layer(o, sd, act, name + ".to_out.0", cfg)


# --------------------------------------------------------------------------- #
# blocks
# --------------------------------------------------------------------------- #
def _gn(x, sd, name, eps):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps)


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def resnet(x: Tensor, temb: Tensor, sd, act, name: str, cfg: QConfig) -> Tensor:
    """QuantResnetBlock2D.forward (quant/quant_block.py:98-119)."""
    h = F.silu(_gn(x, sd, name + ".norm1", 1e-5))
    h = quant_layer(h, sd, act, name + ".conv1", cfg, padding=1)
    t = quant_layer(F.silu(temb), sd, act, name + ".time_emb_proj", cfg)[:, :, None, None]
    h = h + t
    h = F.silu(_gn(h, sd, name + ".norm2", 1e-5))
    h = quant_layer(h, sd, act, name + ".conv2", cfg, padding=1)
    if (name + ".conv_shortcut.w") in sd:
        x = quant_layer(x, sd, act, name + ".conv_shortcut", cfg)
    return x + h


def transformer_block(x, ctx, sd, act, name, cfg, heads) -> Tensor:
    """QuantBasicTransformerBlock.forward (quant/quant_block.py:165-186);
    GEGLU/FeedForward (diffusers_rewrite/sd.py:210-236)."""
    x = attention(_ln(x, sd, name + ".norm1"), None, sd, act, name + ".attn1", cfg,
                  heads=heads, is_cross=False) + x
    x = attention(_ln(x, sd, name + ".norm2"), ctx, sd, act, name + ".attn2", cfg,
                  heads=heads, is_cross=True) + x
    h = quant_layer(_ln(x, sd, name + ".norm3"), sd, act, name + ".ff.net.0.proj", cfg)
    h1, h2 = h.chunk(2, dim=-1)
    h = h1 * F.gelu(h2)
    return quant_layer(h, sd, act, name + ".ff.net.2", cfg) + x


def transformer2d(x, ctx, sd, act, name, cfg, *, n_layers: int, heads: int, linear_proj: bool) -> Tensor:
    """Transformer2DModel.forward (sd.py:283-305: conv 1x1 proj; sdxl.py:306-326: linear proj)."""
    b, c, hh, ww = x.shape
    res = x
    h = _gn(x, sd, name + ".norm", 1e-6)
    if linear_proj:
        h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        h = quant_layer(h, sd, act, name + ".proj_in", cfg)
    else:
        h = quant_layer(h, sd, act, name + ".proj_in", cfg)
        h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, h.shape[1])
    for i in range(n_layers):
        h = transformer_block(h, ctx, sd, act, f"{name}.transformer_blocks.{i}", cfg, heads)
    if linear_proj:
        h = quant_layer(h, sd, act, name + ".proj_out", cfg)
        h = h.reshape(b, hh, ww, -1).permute(0, 3, 1, 2).contiguous()
    else:
        h = h.reshape(b, hh, ww, -1).permute(0, 3, 1, 2).contiguous()
        h = quant_layer(h, sd, act, name + ".proj_out", cfg)
    return h + res


def timestep_embedding(t: Tensor, dim: int) -> Tensor:
    """Timesteps.forward (sd.py:25-39): cos first, then sin."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - 0.0)
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def _time_mlp(x, sd, act, name, cfg):
    h = quant_layer(x, sd, act, name + ".linear_1", cfg)
    return quant_layer(F.silu(h), sd, act, name + ".linear_2", cfg)


# --------------------------------------------------------------------------- #
# UNet topologies (diffusers_rewrite/sd.py:493-544, sdxl.py:505-556)
# --------------------------------------------------------------------------- #
# down block: (in, out, n_transformer_layers or None, has_down, first_shortcut)
SD_SPEC = dict(
    heads=lambda c: 8, ctx_dim=768, linear_proj=False, sample=64,
    down=[(320, 320, 1, True), (320, 640, 1, True), (640, 1280, 1, True), (1280, 1280, None, False)],
    mid_layers=1,
    # up block: (in(skip-most), out, prev, n_layers or None, has_up)
    up=[(1280, 1280, 1280, None, True), (640, 1280, 1280, 1, True), (320, 640, 1280, 1, True),
        (320, 320, 640, 1, False)],
)
SDXL_SPEC = dict(
    heads=lambda c: c // 64, ctx_dim=2048, linear_proj=True, sample=128,
    down=[(320, 320, None, True), (320, 640, 2, True), (640, 1280, 10, False)],
    mid_layers=10,
    up=[(640, 1280, 1280, 10, True), (320, 640, 1280, 2, True), (320, 320, 640, None, False)],
)
SPECS = {"sd": SD_SPEC, "sdxl": SDXL_SPEC}


def unet_forward(model_type: str, sd, act, cfg: QConfig, sample: Tensor, timesteps: Tensor,
                 ctx: Tensor, added: Optional[Dict[str, Tensor]] = None, taps: Optional[list] = None) -> Tensor:
    """UNet2DConditionModel.forward (sd.py:546-620, sdxl.py:558-631) behind
    QuantModel.forward (quant/quant_model.py:113-116)."""
    spec = SPECS[model_type]
    P = "model."
    heads = spec["heads"]
    lp = spec["linear_proj"]
    bsz = sample.shape[0]
    t = timesteps.reshape(-1).expand(bsz)
    emb = _time_mlp(timestep_embedding(t, 320), sd, act, P + "time_embedding", cfg)
    if model_type == "sdxl":
        te = timestep_embedding(added["time_ids"].flatten(), 256).reshape(bsz, -1)
        add = torch.cat([added["text_embeds"], te], dim=-1).to(emb.dtype)
        emb = emb + _time_mlp(add, sd, act, P + "add_embedding", cfg)

    def tap(name, x):
        if taps is not None:
            taps.append((name, x.detach().clone()))

    tap("emb", emb)
    h = quant_layer(sample, sd, act, P + "conv_in", cfg, padding=1, fp_layer=True)
    tap("conv_in", h)
    skips = [h]
    for i, (cin, cout, nl, has_down) in enumerate(spec["down"]):
        for j in range(2):
            h = resnet(h, emb, sd, act, f"{P}down_blocks.{i}.resnets.{j}", cfg)
            tap(f"down{i}.res{j}", h)
            if nl is not None:
                h = transformer2d(h, ctx, sd, act, f"{P}down_blocks.{i}.attentions.{j}", cfg,
                                  n_layers=nl, heads=heads(cout), linear_proj=lp)
                tap(f"down{i}.attn{j}", h)
            skips.append(h)
        if has_down:
            h = quant_layer(h, sd, act, f"{P}down_blocks.{i}.downsamplers.0.conv", cfg, stride=2, padding=1)
            skips.append(h)

    h = resnet(h, emb, sd, act, P + "mid_block.resnets.0", cfg)
    h = transformer2d(h, ctx, sd, act, P + "mid_block.attentions.0", cfg,
                      n_layers=spec["mid_layers"], heads=heads(1280), linear_proj=lp)
    h = resnet(h, emb, sd, act, P + "mid_block.resnets.1", cfg)
    tap("mid", h)

    for i, (cin, cout, prev, nl, has_up) in enumerate(spec["up"]):
        for j in range(3):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet(h, emb, sd, act, f"{P}up_blocks.{i}.resnets.{j}", cfg)
            tap(f"up{i}.res{j}", h)
            if nl is not None:
                h = transformer2d(h, ctx, sd, act, f"{P}up_blocks.{i}.attentions.{j}", cfg,
                                  n_layers=nl, heads=heads(cout), linear_proj=lp)
                tap(f"up{i}.attn{j}", h)
        if has_up:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = quant_layer(h, sd, act, f"{P}up_blocks.{i}.upsamplers.0.conv", cfg, padding=1)

    h = F.silu(_gn(h, sd, P + "conv_norm_out", 1e-5))
    return quant_layer(h, sd, act, P + "conv_out", cfg, padding=1, fp_layer=True)


# --------------------------------------------------------------------------- #
# time-aware step selection (quant/calibration.py:297-312)
# --------------------------------------------------------------------------- #
def act_index(t: float, num_inference_steps: int) -> int:
    return int((1000 - t) // (1000 // num_inference_steps))


def update_group_convs(cfg: QConfig, act: Dict[str, Tensor], sd: Dict[str, Tensor]) -> None:
    """Sticky use_group_num flip (quant/calibration.py:271-278): a conv layer
    switches to the unfold path the first time its checkpoint delta is not the
    scalar the dummy forward produced."""
    for k, v in act.items():
        if k.endswith(".aqtizer.delta") and v.dim() > 0:
            name = k[: -len(".aqtizer.delta")]
            if sd[name + ".w"].dim() == 4:
                cfg.group_convs.add(name)
