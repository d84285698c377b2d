"""T2ILogQuantizer with the reference's interface (quant/quant_layer_text.py), CUDA-executed.

Stand-alone `forward` runs dgq_t2i_log_quant_f32 (+ dgq_max_f32 for real-time delta); inside the
UNet the same arithmetic is fused into the attention kernel (dgq_attention, DGQ_MAP_LOG2) and this
module only carries the configuration (level, real_time, delta)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .quant_layer import Scaler, _need_cuda


class T2ILogQuantizer(nn.Module):
    def __init__(self, bits: int = 8, symmetric: bool = False, channel_wise: bool = False,
                 scaler=Scaler.MINMAX, leaf_param: bool = False, always_zero: bool = True,
                 quant_emb: bool = False, real_time: bool = False, log_max_1: bool = False) -> None:
        super().__init__()
        self.level = 2 ** bits
        self.symmetric = symmetric
        self.channel_wise = channel_wise
        self.scaler = scaler
        self.leaf_param = leaf_param
        self.running_stat = False
        self.always_zero = always_zero
        self.delta = None
        self.zero_point = None
        self.init = False
        self.quant_emb = quant_emb
        self.real_time = real_time
        self.NB, self.PB = 0, self.level - 1
        self.log_max_1 = log_max_1

    def static_delta(self, device) -> torch.Tensor:
        """Device scalar used when real_time is off.  The reference never saves this delta (its
        state dict fails the 2-key filter, calibration_group_quantization.py:104), so at inference it
        is whatever the loader's random dummy forward left behind.  Here it must be
        set explicitly (`q.delta = tensor`) or log_max_1 must be on."""
        if self.log_max_1:
            return torch.ones(1, dtype=torch.float32, device=device)
        if self.delta is None:
            raise RuntimeError("T2ILogQuantizer without real_time needs an explicit delta: the reference "
                               "does not store one in its checkpoints (see DESIGN.md, quirk H6-i)")
        return self.delta.detach().to(device=device, dtype=torch.float32).reshape(1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _need_cuda(x, "T2ILogQuantizer.forward")
        xf = x.detach().to(torch.float32).contiguous()
        delta = None if self.real_time else self.static_delta(x.device)
        return ops.t2i_log_quant(xf, delta, float(self.level - 1)).to(x.dtype)

    def bitwidth_refactor(self, bits: int = 8) -> None:
        self.level = 2 ** bits

    def half(self):
        return self

    def float(self):
        return self
