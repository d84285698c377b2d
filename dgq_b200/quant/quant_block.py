"""Quant blocks with the reference's interface (quant/quant_block.py): they re-host the
sub-modules of a ResnetBlock2D / BasicTransformerBlock under the same attribute names, attach the
attention quantizers (aqtizer_q/k/v/w, start_peak) and toggle quantization state.  `forward`
runs dgq_b200.engine's fused kernels."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from .. import engine, ops
from ..unet.common import Attention  # the class `isinstance` checks bind to (reference quant_block.py:3-4,32)
from .quant_layer import QuantLayer, UniformAffineQuantizer, StraightThrough
from .quant_layer_text import T2ILogQuantizer


class BaseQuantBlock(nn.Module):
    def __init__(self, aq_params: dict = {}) -> None:
        super().__init__()
        self.use_wq = False
        self.use_aq = False
        self.act_func = StraightThrough()
        self.ignore_recon = False

    def set_quant_state(self, use_wq: bool = False, use_aq: bool = False) -> None:
        for m in self.modules():
            if isinstance(m, QuantLayer):
                m.set_quant_state(use_wq=use_wq, use_aq=use_aq)
            if isinstance(m, Attention):
                m.use_aq = use_aq


class QuantResnetBlock2D(BaseQuantBlock):
    def __init__(self, resnet, aq_params: dict = {}) -> None:
        super().__init__(aq_params)
        self.norm1 = resnet.norm1
        self.conv1 = resnet.conv1
        self.time_emb_proj = resnet.time_emb_proj
        self.norm2 = resnet.norm2
        self.dropout = resnet.dropout
        self.conv2 = resnet.conv2
        self.nonlinearity = resnet.nonlinearity
        self.conv_shortcut = resnet.conv_shortcut

    def forward(self, input_tensor, temb):
        x = engine.act_from_nchw(input_tensor)
        silu_emb = ops.silu(temb.detach().to(ops.ACT_DTYPE).contiguous())
        return engine.act_to_nchw(engine.resnet(self, x, silu_emb), dtype=input_tensor.dtype)


class QuantBasicTransformerBlock(BaseQuantBlock):
    def __init__(self, tran, aq_params: dict = {}, softmax_aq_params: dict = {}) -> None:
        super().__init__(aq_params)
        self.norm1 = tran.norm1
        self.attn1 = tran.attn1
        self.norm2 = tran.norm2
        self.attn2 = tran.attn2
        self.norm3 = tran.norm3
        self.ff = tran.ff
        for attn in (self.attn1, self.attn2):
            attn.aqtizer_q = UniformAffineQuantizer(**aq_params)
            attn.aqtizer_k = UniformAffineQuantizer(**aq_params)
            attn.aqtizer_v = UniformAffineQuantizer(**aq_params)
        aq_params_w = dict(aq_params)
        aq_params_w["bits"] = softmax_aq_params["softmax_a_bit"]
        aq_params_w["symmetric"] = False
        aq_params_w["always_zero"] = True
        if softmax_aq_params["t2i_log_quant"]:
            aq_params_w["real_time"] = softmax_aq_params["t2i_real_time"]
            aq_params_w["log_max_1"] = softmax_aq_params["log_max_1"]
            self.attn1.aqtizer_w = T2ILogQuantizer(**aq_params_w)
            self.attn2.aqtizer_w = T2ILogQuantizer(**aq_params_w)
        else:
            self.attn1.aqtizer_w = UniformAffineQuantizer(**aq_params_w)
            self.attn2.aqtizer_w = UniformAffineQuantizer(**aq_params_w)
        if softmax_aq_params["t2i_start_peak"]:
            self.attn2.start_peak = True  # cross-attention only (reference :157-158)
        self.attn1.use_aq = False
        self.attn2.use_aq = False

    def forward(self, x, encoder_hidden_states=None):
        out = engine.transformer_block(self, engine.act_from_tokens(x), encoder_hidden_states)
        return engine.act_to_tokens(out, x.dtype)


def b2qb() -> Dict[str, type]:
    return {"ResnetBlock2D": QuantResnetBlock2D, "BasicTransformerBlock": QuantBasicTransformerBlock}
