"""UNet building blocks shared by the SD-v1.4 and SDXL graphs.

Module and attribute names follow the reference's hand-flattened graphs
(diffusers_rewrite/sd.py, sdxl.py) so that `QuantModel.state_dict()` keys, `quant_module` /
`quant_block` tree surgery and user code keep working.  The modules only own parameters and
structure: every `forward` converts its torch-layout arguments to the engine's NHWC fp16
activations and calls dgq_b200.engine, which runs the CUDA kernels.  A graph whose Conv2d/Linear
layers have not been wrapped by `QuantModel` (plain nn layers) is outside this path and raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import engine, ops


def _check_wrapped(layer, owner: str) -> None:
    if isinstance(layer, (nn.Conv2d, nn.Linear)):
        raise NotImplementedError(
            f"{owner}: this UNet runs only as a quantized model -- wrap it with quant.quant_model.QuantModel "
            "(set_quant_state(False, False) gives the un-quantized fp16 path)")


class Timesteps(nn.Module):
    def __init__(self, num_channels: int = 320):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, timesteps):
        return ops.timestep_embedding(timesteps.float(), self.num_channels, f32=True)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, out_features, bias=True)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(out_features, out_features, bias=True)

    def forward(self, sample):
        _check_wrapped(self.linear_1, "TimestepEmbedding")
        x = sample.detach()
        x = x.contiguous() if x.dtype in (torch.float32, torch.float16) else x.float().contiguous()
        return engine.time_mlp(self, x).to(sample.dtype)


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, conv_shortcut=True):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=1e-05, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(1280, out_channels, bias=True)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=1e-05, affine=True)
        self.dropout = nn.Dropout(p=0.0, inplace=False)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = None
        if conv_shortcut:
            self.conv_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1)

    def forward(self, input_tensor, temb):
        _check_wrapped(self.conv1, "ResnetBlock2D")
        x = engine.act_from_nchw(input_tensor)
        silu_emb = ops.silu(temb.detach().to(ops.ACT_DTYPE).contiguous())
        return engine.act_to_nchw(engine.resnet(self, x, silu_emb), dtype=input_tensor.dtype)


class Attention(nn.Module):
    def __init__(self, inner_dim, cross_attention_dim=None, num_heads=None, dropout=0.0):
        super().__init__()
        if num_heads is None:
            self.head_dim = 64
            self.num_heads = inner_dim // self.head_dim
        else:
            self.num_heads = num_heads
            self.head_dim = inner_dim // num_heads
        self.scale = self.head_dim ** -0.5
        if cross_attention_dim is None:
            cross_attention_dim = inner_dim
        self.to_q = nn.Linear(inner_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(cross_attention_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(cross_attention_dim, inner_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner_dim, inner_dim), nn.Dropout(dropout, inplace=False)])

    def Attention_forward(self, hidden_states, encoder_hidden_states=None):
        """Quantization-aware attention (reference sd.py:151-207); also serves `forward`: with
        use_aq off it is the plain softmax attention of sd.py:122-149."""
        _check_wrapped(self.to_q, "Attention")
        dev = hidden_states.device
        b, t, c = hidden_states.shape
        x = hidden_states.detach().reshape(b * t, c)
        x = x.contiguous() if x.dtype in (torch.float32, torch.float16) else x.float().contiguous()
        qs = [self.to_q.act_qparam(dev), self.to_k.act_qparam(dev), self.to_v.act_qparam(dev)]
        if encoder_hidden_states is not None:
            cx, _, s = engine._ctx_operand(encoder_hidden_states)
            xq = ops.row_quant(x, qs[:1], emit_int=[engine._emit(self.to_q, qs[0])])[0]
            xk, xv = ops.row_quant(cx, qs[1:], emit_int=[engine._emit(self.to_k, qs[1]), engine._emit(self.to_v, qs[2])])
        else:
            s = t
            xq, xk, xv = ops.row_quant(x, qs, emit_int=[engine._emit(l, q) for l, q in zip((self.to_q, self.to_k, self.to_v), qs)])
        out = engine.attention(self, xq, xk, xv, qs, b, t, s, resid=None)
        return out.view(b, t, -1).to(hidden_states.dtype)

    def forward(self, hidden_states, encoder_hidden_states=None):
        return self.Attention_forward(hidden_states, encoder_hidden_states)


class GEGLU(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.proj = nn.Linear(in_features, out_features * 2, bias=True)

    def forward(self, x):
        _check_wrapped(self.proj, "GEGLU")
        shp = x.shape
        x2 = x.detach().reshape(-1, shp[-1])
        x2 = x2.contiguous() if x2.dtype in (torch.float32, torch.float16) else x2.float().contiguous()
        g = engine.linear(self.proj, *engine.quant_rows(x2, self.proj))
        return ops.geglu_quant(g, ops.NOQ).view(*shp[:-1], -1).to(x.dtype)


class FeedForward(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(in_features, out_features * 4), nn.Dropout(p=0.0, inplace=False),
                                  nn.Linear(out_features * 4, out_features, bias=True)])

    def forward(self, x):
        for layer in self.net:
            x = layer(x)
        return x


class BasicTransformerBlockBase(nn.Module):
    def __init__(self, hidden_size, cross_dim, num_heads=None):
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden_size, eps=1e-05, elementwise_affine=True)
        self.attn1 = Attention(hidden_size, num_heads=num_heads)
        self.norm2 = nn.LayerNorm(hidden_size, eps=1e-05, elementwise_affine=True)
        self.attn2 = Attention(hidden_size, cross_dim, num_heads=num_heads)
        self.norm3 = nn.LayerNorm(hidden_size, eps=1e-05, elementwise_affine=True)
        self.ff = FeedForward(hidden_size, hidden_size)

    def forward(self, x, encoder_hidden_states=None):
        _check_wrapped(self.attn1.to_q, "BasicTransformerBlock")
        out = engine.transformer_block(self, engine.act_from_tokens(x), encoder_hidden_states)
        return engine.act_to_tokens(out, x.dtype)


class Transformer2DModelBase(nn.Module):
    def forward(self, hidden_states, encoder_hidden_states=None):
        _check_wrapped(self.proj_in, "Transformer2DModel")
        out = engine.transformer2d(self, engine.act_from_nchw(hidden_states), encoder_hidden_states)
        return engine.act_to_nchw(out, dtype=hidden_states.dtype)


class Downsample2D(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=2, padding=1)

    def forward(self, x):
        _check_wrapped(self.conv, "Downsample2D")
        return engine.act_to_nchw(engine.conv(self.conv, engine.act_from_nchw(x)), dtype=x.dtype)


class Upsample2D(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        _check_wrapped(self.conv, "Upsample2D")
        return engine.act_to_nchw(engine.conv(self.conv, engine.act_from_nchw(x), upsample=True), dtype=x.dtype)


class _BlockList(nn.Module):
    """Down/Up/Mid containers: structure only.  They are executed by engine.unet_forward, which
    fuses the skip concatenation and up-sampling into the consumers' producer kernels."""

    def forward(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is executed as part of UNet2DConditionModel.forward "
                                  "(skip concat / upsample are fused across block boundaries)")


class UNetConfig:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class UNetBase(nn.Module):
    """Plain nn.Module with the `.config` / `.device` / `.dtype` surface the diffusers pipelines read
    (the reference derives from diffusers' ModelMixin/ConfigMixin, which is not installed here)."""

    def register_to_config(self, **kw):
        self.config = UNetConfig(**kw)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype
