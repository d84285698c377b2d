"""VAE decode on the device -- the step that follows the denoising loop in both reference pipelines
(SURVEY.md 8f row 4):

    image = vae.decode(latents / vae.config.scaling_factor, return_dict=False)[0]
    image = image_processor.postprocess(image, output_type="pil")

(diffusers/src/diffusers/pipelines/stable_diffusion/pipeline_stable_diffusion.py:1066-1069,
pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl.py:1295-1307).  The decoder is not quantized by DGQ; it
runs here on the same kernels as the UNet's full-precision layers: NHWC fp32 activations between kernels, GroupNorm
statistics + (GroupNorm, SiLU, nearest-2x upsample, im2col) producers writing fp16 A operands, `dgq_gemm_f16` with
bias / residual epilogues.  The single-head, 512-channel mid-block attention (attention_processor.py:1200-1262) does
not fit the TMEM-resident flash kernel (head dim <= 256) and runs once per image, so it is three plain GEMMs around
a row-softmax kernel: S = q k^T (fp32), P = softmax(S / sqrt(c)) (fp16), O = P v with v^T produced directly as
W_v x^T and the v bias added after the product (softmax rows sum to 1).

`VaeDecoder` keeps the reference's module names, so a diffusers `AutoencoderKL` state dict loads with
`load_state_dict(..., strict=False)` (only `post_quant_conv.*` and `decoder.*` are used).  No CPU fallback.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from . import engine, ops
from .engine import Act
from .quant.quant_layer import QuantLayer

GN_GROUPS, GN_EPS = 32, 1e-6   # models/autoencoders/vae.py:236-271 (resnet_eps=1e-6, norm_num_groups=32)


class _Resnet(nn.Module):      # ResnetBlock2D with temb_channels=None
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(GN_GROUPS, cin, eps=GN_EPS)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(GN_GROUPS, cout, eps=GN_EPS)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)


class _Attention(nn.Module):   # Attention(heads=1, dim_head=c, bias=True, residual_connection=True, norm_num_groups=32)
    def __init__(self, c: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(GN_GROUPS, c, eps=GN_EPS)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])


class _Mid(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.attentions = nn.ModuleList([_Attention(c)])
        self.resnets = nn.ModuleList([_Resnet(c, c), _Resnet(c, c)])


class _Upsample(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)


class _UpBlock(nn.Module):     # UpDecoderBlock2D
    def __init__(self, cin: int, cout: int, layers: int, upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if j == 0 else cout, cout) for j in range(layers)])
        if upsample:
            self.upsamplers = nn.ModuleList([_Upsample(cout)])


class _Decoder(nn.Module):
    def __init__(self, latent_channels, block_out_channels, layers_per_block):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(latent_channels, rev[0], 3, padding=1)
        self.mid_block = _Mid(rev[0])
        ups, cout = [], rev[0]
        for i, c in enumerate(rev):
            cin, cout = cout, c
            ups.append(_UpBlock(cin, cout, layers_per_block + 1, i != len(rev) - 1))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(GN_GROUPS, block_out_channels[0], eps=GN_EPS)
        self.conv_out = nn.Conv2d(block_out_channels[0], 3, 3, padding=1)


class VaeDecoder(nn.Module):
    """`post_quant_conv` + `decoder` of the reference's AutoencoderKL (autoencoder_kl.py:104-111, 268-279)."""

    # im2col bytes one decode chunk may hold: the widest layer of a chunk is (pixels x 9 x 256 channels) fp16
    CHUNK_BYTES = 12 << 30

    def __init__(self, block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2, latent_channels: int = 4,
                 scaling_factor: float = 0.18215):
        super().__init__()
        self.config = SimpleNamespace(block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      latent_channels=latent_channels, scaling_factor=scaling_factor,
                                      force_upcast=False)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.decoder = _Decoder(latent_channels, block_out_channels, layers_per_block)
        self._fp: dict = {}

    # the kernels take their operands from QuantLayer's full-precision path (weights as fp16, no quantizers: what the
    # UNet's conv_in / conv_out use under disable_out_quantization); wrappers are made on first use and dropped
    # whenever the parameters may have changed
    def _ql(self, layer: nn.Module) -> QuantLayer:
        hit = self._fp.get(id(layer))
        key = (layer.weight._version, layer.weight.data_ptr())
        if hit is None or hit[0] != key:
            ql = QuantLayer(layer, {"bits": 8, "channel_wise": True}, {"bits": 8, "channel_wise": False})
            ql.set_quant_state(False, False)
            hit = (key, ql)
            self._fp[id(layer)] = hit
        return hit[1]

    def _apply(self, fn, *a, **k):
        self._fp = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._fp = {}
        mine = {k: v for k, v in state_dict.items() if k.startswith(("decoder.", "post_quant_conv."))}
        return super().load_state_dict(mine if not strict else state_dict, strict=strict, **kw)

    # ---- blocks ---------------------------------------------------------------------------------------------
    def _conv(self, layer, x: Act, **kw) -> Act:
        return engine.conv(self._ql(layer), x, **kw)

    def _resnet(self, blk: _Resnet, x: Act) -> Act:
        sc = self._conv(blk.conv_shortcut, x).t if hasattr(blk, "conv_shortcut") else x.t
        h = self._conv(blk.conv1, x, gn=engine._gn(blk.norm1, x), act=1)
        return self._conv(blk.conv2, h, gn=engine._gn(blk.norm2, h), act=1, resid=sc)

    def _attention(self, attn: _Attention, x: Act) -> Act:
        c, t = x.c, x.rows
        # GroupNorm over (c, hw) without an activation, as the fp16 operand of the three projections
        xn = ops.act_producer(x.t, batch=x.b, h=x.h, w=x.w, ksize=1, gn=engine._gn(attn.group_norm, x), act=0)
        q = engine._gemm(self._ql(attn.to_q), xn, want_f32=False)
        k = engine._gemm(self._ql(attn.to_k), xn, want_f32=False)
        wv, _, bv, _ = self._ql(attn.to_v).packed()
        o = torch.empty(x.b * t, c, dtype=torch.float16, device=x.t.device)
        s = torch.empty(t, t, dtype=torch.float32, device=x.t.device)
        p = torch.empty(t, t, dtype=torch.float16, device=x.t.device)
        for i in range(x.b):
            rows = slice(i * t, (i + 1) * t)
            vt = ops.gemm(wv, xn[rows], t)                               # v^T = W_v x^T  [c, t]
            ops.gemm(q[rows], k[rows], t, out=s)                         # S = q k^T
            ops.softmax_rows(s, 1.0 / math.sqrt(c), out=p)
            ops.gemm(p, vt, c, bias=bv, out=o[rows])                     # O = P v + b_v
        out = engine._gemm(self._ql(attn.to_out[0]), o, resid=x.t)
        return Act(out, x.b, x.h, x.w)

    def _decode_chunk(self, z: torch.Tensor) -> torch.Tensor:
        d = self.decoder
        x = self._conv(self.post_quant_conv, engine.act_from_nchw(z))
        x = self._conv(d.conv_in, x)
        x = self._resnet(d.mid_block.resnets[0], x)
        x = self._attention(d.mid_block.attentions[0], x)
        x = self._resnet(d.mid_block.resnets[1], x)
        for blk in d.up_blocks:
            for r in blk.resnets:
                x = self._resnet(r, x)
            if hasattr(blk, "upsamplers"):
                x = self._conv(blk.upsamplers[0].conv, x, upsample=True)
        x = self._conv(d.conv_out, x, gn=engine._gn(d.conv_norm_out, x), act=1)
        return engine.act_to_nchw(x, c=3)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = False):
        """AutoencoderKL.decode: z [b, latent_channels, h, w] (already divided by scaling_factor) -> ([b, 3, 8h, 8w],)"""
        if not z.is_cuda:
            raise RuntimeError("dgq_b200.vae: CUDA tensors only (there is no CPU path)")
        if return_dict:
            raise NotImplementedError("return_dict=True: the reference pipelines call decode(..., return_dict=False)")
        b, _, h, w = z.shape
        up = 2 ** (len(self.config.block_out_channels) - 1)
        per_image = h * up * w * up * 9 * 256 * 2
        step = max(1, min(b, self.CHUNK_BYTES // per_image))
        outs = [self._decode_chunk(z[i:i + step].float().contiguous()) for i in range(0, b, step)]
        return (outs[0] if len(outs) == 1 else torch.cat(outs, 0),)

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """the pipelines' call site: decode(latents / scaling_factor)[0]"""
        return self.decode(latents / self.config.scaling_factor)[0]


def postprocess(image: torch.Tensor) -> torch.Tensor:
    """VaeImageProcessor.postprocess up to the PIL conversion (image_processor.py:138-142, 84-97):
    [b, 3, H, W] in [-1, 1] -> uint8 [b, H, W, 3] on the host"""
    x = (image / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).float().cpu().numpy()
    return torch.from_numpy((x * 255).round().astype("uint8"))


def save_images(u8: torch.Tensor, outdir: str, prefix: str = "img") -> list:
    """uint8 [b, H, W, 3] -> PNG files (PIL), like the reference's `images[i].save(...)` (src/inference_qmodel.py:100-108)"""
    import os
    from PIL import Image
    os.makedirs(outdir, exist_ok=True)
    paths = []
    for i in range(u8.shape[0]):
        path = os.path.join(outdir, f"{prefix}_{i:04d}.png")
        Image.fromarray(u8[i].numpy()).save(path)
        paths.append(path)
    return paths
