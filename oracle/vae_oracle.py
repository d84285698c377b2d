"""CPU fp32 restatement of the VAE decode that follows the denoising loop -- TEST INFRASTRUCTURE ONLY
(imported by tests/, tests/golden/make_golden.py and bench legs; never by dgq_b200/).

Restates, functionally and keyed by the reference's own state-dict names:

* ``AutoencoderKL._decode``: ``post_quant_conv`` then ``Decoder``
  (diffusers/src/diffusers/models/autoencoders/autoencoder_kl.py:268-279);
* ``Decoder.forward``: conv_in -> mid_block -> up_blocks -> conv_norm_out -> SiLU -> conv_out
  (models/autoencoders/vae.py:277-340);
* ``UNetMidBlock2D`` (resnet, attention, resnet; models/unets/unet_2d_blocks.py:514-666) with the
  single-head ``Attention`` block: GroupNorm over (b, c, hw), q/k/v linears with bias,
  softmax(q k^T / sqrt(c)) v, to_out, + residual (models/attention_processor.py:1200-1262);
* ``UpDecoderBlock2D`` (resnets, then nearest-2x ``Upsample2D`` + 3x3 conv;
  unet_2d_blocks.py:2549-2650, models/upsampling.py:160-186);
* ``ResnetBlock2D`` without a time embedding (GN -> SiLU -> conv -> GN -> SiLU -> conv, 1x1 shortcut when the
  channel count changes; models/resnet.py);
* the pipelines' call sites: ``vae.decode(latents / scaling_factor)``
  (pipelines/stable_diffusion/pipeline_stable_diffusion.py:1066-1069,
  pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl.py:1295-1307) and
  ``VaeImageProcessor.postprocess`` (image_processor.py:138-142, 84-97).

Pinned by tests/golden/vae.pt: outputs of the reference's AutoencoderKL itself on weights synthesised here
(``make_vae_state``: numpy PCG64 per tensor name, machine independent), minted by
``python tests/golden/make_golden.py vae``.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, Iterator, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# the two decoders the reference's pipelines load (stabilityai sd-vae / sdxl-vae share this architecture)
VAE_CONFIGS = {
    "sd": dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, scaling_factor=0.18215),
    "sdxl": dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, scaling_factor=0.13025),
    # reduced width / depth, same block structure: the golden fixture that runs in seconds on a CPU
    "small": dict(block_out_channels=(128, 128, 256, 256), layers_per_block=1, latent_channels=4, scaling_factor=0.18215),
}
GN_GROUPS, GN_EPS = 32, 1e-6


def iter_decoder_modules(cfg: dict) -> Iterator[Tuple[str, tuple]]:
    """(name, descriptor) of every parametrised module of post_quant_conv + decoder, in forward order."""
    boc, lpb, lc = cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"]
    yield "post_quant_conv", ("conv", lc, lc, 1)
    top = boc[-1]
    yield "decoder.conv_in", ("conv", top, lc, 3)

    def resnet(p, cin, cout):
        yield p + ".norm1", ("gn", cin)
        yield p + ".conv1", ("conv", cout, cin, 3)
        yield p + ".norm2", ("gn", cout)
        yield p + ".conv2", ("conv", cout, cout, 3)
        if cin != cout:
            yield p + ".conv_shortcut", ("conv", cout, cin, 1)

    yield from resnet("decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    yield a + ".group_norm", ("gn", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        yield f"{a}.{n}", ("lin", top, top)
    yield from resnet("decoder.mid_block.resnets.1", top, top)
    rev = list(reversed(boc))
    cout = rev[0]
    for i, c in enumerate(rev):
        cin, cout = cout, c
        for j in range(lpb + 1):
            yield from resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i != len(rev) - 1:
            yield f"decoder.up_blocks.{i}.upsamplers.0.conv", ("conv", cout, cout, 3)
    yield "decoder.conv_norm_out", ("gn", boc[0])
    yield "decoder.conv_out", ("conv", 3, boc[0], 3)


def _rng(seed: int, name: str) -> np.random.Generator:
    h = hashlib.sha256(f"vae:{seed}:{name}".encode()).digest()
    return np.random.Generator(np.random.PCG64(int.from_bytes(h[:8], "little")))


def make_vae_state(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Random-init decoder weights (uniform +-1/sqrt(fan_in) like nn.Conv2d / nn.Linear, norm affine parameters
    perturbed off (1, 0)), rounded to fp16-representable values so that an fp16 operand holds them exactly."""
    sd: Dict[str, Tensor] = {}
    for name, d in iter_decoder_modules(cfg):
        g = _rng(seed, name)
        if d[0] == "gn":
            sd[name + ".weight"] = torch.from_numpy((1.0 + 0.1 * g.standard_normal(d[1])).astype(np.float32))
            sd[name + ".bias"] = torch.from_numpy((0.1 * g.standard_normal(d[1])).astype(np.float32))
            continue
        if d[0] == "conv":
            _, co, ci, k = d
            shape, fan_in = (co, ci, k, k), ci * k * k
        else:
            _, co, ci = d
            shape, fan_in = (co, ci), ci
        bound = 1.0 / math.sqrt(fan_in)
        w = torch.from_numpy(g.uniform(-bound, bound, size=shape).astype(np.float32))
        sd[name + ".weight"] = w.half().float()
        sd[name + ".bias"] = torch.from_numpy(g.uniform(-bound, bound, size=(co,)).astype(np.float32))
    return sd


# --------------------------------------------------------------------------- #
def _gn(sd, p, x):
    return F.group_norm(x, GN_GROUPS, sd[p + ".weight"], sd[p + ".bias"], GN_EPS)


def _conv(sd, p, x):
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd[p + ".bias"], padding=w.shape[-1] // 2)


def resnet(sd, p, x):
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h)))
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(sd, p + ".conv_shortcut", x)
    return x + h


def attention(sd, p, x):
    b, c, hh, ww = x.shape
    t = _gn(sd, p + ".group_norm", x.view(b, c, hh * ww)).transpose(1, 2)          # [b, hw, c]
    q = F.linear(t, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(t, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(t, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1) @ v
    o = F.linear(a, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, hh, ww)


def decode(sd: Dict[str, Tensor], cfg: dict, z: Tensor, taps: dict = None) -> Tensor:
    """AutoencoderKL.decode(z).sample: z [b, latent_channels, h, w] -> [b, 3, 8h, 8w]"""
    n_up = len(cfg["block_out_channels"])
    x = _conv(sd, "post_quant_conv", z.float())
    x = _conv(sd, "decoder.conv_in", x)
    x = resnet(sd, "decoder.mid_block.resnets.0", x)
    x = attention(sd, "decoder.mid_block.attentions.0", x)
    x = resnet(sd, "decoder.mid_block.resnets.1", x)
    if taps is not None:
        taps["mid"] = x
    for i in range(n_up):
        for j in range(cfg["layers_per_block"] + 1):
            x = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i != n_up - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
        if taps is not None:
            taps[f"up{i}"] = x
    x = F.silu(_gn(sd, "decoder.conv_norm_out", x))
    return _conv(sd, "decoder.conv_out", x)


def decode_latents(sd, cfg, latents: Tensor) -> Tensor:
    """the pipelines' call: vae.decode(latents / vae.config.scaling_factor)"""
    return decode(sd, cfg, latents / cfg["scaling_factor"])


def postprocess(image: Tensor) -> Tensor:
    """VaeImageProcessor.postprocess(output_type='pil') up to the PIL object: [b, 3, H, W] in [-1, 1] ->
    uint8 [b, H, W, 3]"""
    x = (image / 2 + 0.5).clamp(0, 1)
    return torch.from_numpy((x.permute(0, 2, 3, 1).float().numpy() * 255).round().astype("uint8"))
