"""Teacher-forcing harness: run the ORACLE's UNet forward and, at every QuantLayer / attention / resnet /
transformer block, hand the oracle's own INPUT (and its output) to a callback.  The callback runs the CUDA module
of the same name on that input, so every comparison starts from bit-identical inputs and the chaotic
error growth of a random-init quantized UNet (tests/golden/self_sensitivity.json) never enters.

Test infrastructure only (it monkey-patches oracle.dgq_oracle's module globals for the duration of the `with`)."""
from __future__ import annotations

import torch

from oracle import dgq_oracle as O


class LayerTrace:
    """with LayerTrace(cb): O.unet_forward(...)   ->   cb(kind, name, inputs: dict, out: Tensor)

    kind in {"quant_layer", "attention_core", "attention", "resnet", "transformer_block"}."""

    def __init__(self, cb, kinds=("quant_layer", "attention_core", "attention", "resnet", "transformer_block")):
        self.cb = cb
        self.kinds = set(kinds)

    def __enter__(self):
        self._orig = {n: getattr(O, n) for n in ("quant_layer", "attention_core", "attention", "resnet",
                                                 "transformer_block")}
        cb, kinds, orig = self.cb, self.kinds, self._orig

        def quant_layer(x, sd, act, name, cfg, *, stride=1, padding=0, fp_layer=False):
            y = orig["quant_layer"](x, sd, act, name, cfg, stride=stride, padding=padding, fp_layer=fp_layer)
            if "quant_layer" in kinds:
                cb("quant_layer", name, dict(x=x, stride=stride, padding=padding, fp_layer=fp_layer, sd=sd, act=act,
                                             cfg=cfg), y)
            return y

        def attention_core(q, k, v, act, name, cfg, *, is_cross, return_probs=False):
            o = orig["attention_core"](q, k, v, act, name, cfg, is_cross=is_cross, return_probs=return_probs)
            if "attention_core" in kinds and not return_probs:
                cb("attention_core", name, dict(q=q, k=k, v=v, act=act, cfg=cfg, is_cross=is_cross), o)
            return o

        def attention(x, ctx, sd, act, name, cfg, *, heads, is_cross):
            o = orig["attention"](x, ctx, sd, act, name, cfg, heads=heads, is_cross=is_cross)
            if "attention" in kinds:
                cb("attention", name, dict(x=x, ctx=ctx, heads=heads, is_cross=is_cross), o)
            return o

        def resnet(x, temb, sd, act, name, cfg):
            o = orig["resnet"](x, temb, sd, act, name, cfg)
            if "resnet" in kinds:
                cb("resnet", name, dict(x=x, temb=temb), o)
            return o

        def transformer_block(x, ctx, sd, act, name, cfg, heads):
            o = orig["transformer_block"](x, ctx, sd, act, name, cfg, heads)
            if "transformer_block" in kinds:
                cb("transformer_block", name, dict(x=x, ctx=ctx, heads=heads), o)
            return o

        O.quant_layer, O.attention_core, O.attention = quant_layer, attention_core, attention
        O.resnet, O.transformer_block = resnet, transformer_block
        return self

    def __exit__(self, *exc):
        for n, f in self._orig.items():
            setattr(O, n, f)
        return False


def max_rel(y: torch.Tensor, ref: torch.Tensor) -> float:
    """max |y - ref| / max |ref|  (the per-layer tolerance of BASELINE.json north_star)"""
    return ((y.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-30)).item()


def rel_l2(y: torch.Tensor, ref: torch.Tensor) -> float:
    return ((y.float() - ref.float()).norm() / ref.float().norm().clamp_min(1e-30)).item()
