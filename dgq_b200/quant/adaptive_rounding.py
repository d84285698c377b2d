"""AdaRoundQuantizer -- inference (hard-rounding) half only.

A BRECQ checkpoint carries `wqtizer.alpha`; the reference converts every weight quantizer with
uaq2adar (quant/calibration.py:20-43, 227-230) and then re-rounds the weights on EVERY forward
(quant/adaptive_rounding.py:51-70).  Here the module only holds (delta, zero_point, alpha); the
rounding `floor(w/delta) + (alpha >= 0)` happens once inside dgq_pack_weight."""
from __future__ import annotations

from enum import Enum

import torch
import torch.nn as nn

from .. import ops
from .quant_layer import UniformAffineQuantizer, _need_cuda

RMODE = Enum("RMODE", ("LEARNED_ROUND_SIGMOID", "NEAREST", "NEAREST_STE", "STOCHASTIC", "LEARNED_HARD_SIGMOID"))


class AdaRoundQuantizer(nn.Module):
    def __init__(self, uaqtizer: UniformAffineQuantizer, w: torch.Tensor,
                 rmode: RMODE = RMODE.LEARNED_HARD_SIGMOID) -> None:
        super().__init__()
        if rmode != RMODE.LEARNED_HARD_SIGMOID:
            raise NotImplementedError("only the learned-hard-sigmoid AdaRound mode reaches inference")
        self.level = uaqtizer.level
        self.symmetric = uaqtizer.symmetric
        self.delta = uaqtizer.delta
        self.zero_point = uaqtizer.zero_point
        self.rmode = rmode
        self.soft_tgt = False
        self.gamma, self.zeta = -0.1, 1.1
        # alpha is overwritten by the checkpoint; the reference's init (adaptive_rounding.py:32-39)
        # reproduces nearest rounding, i.e. alpha >= 0 exactly where frac(w/delta) >= 0.5
        d = self.delta.detach().to(w.device)
        rest = (w / d) - torch.floor(w / d)
        self.alpha = nn.Parameter(-torch.log((self.zeta - self.gamma) / (rest - self.gamma) - 1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """De-quantised weights with hard rounding (reference :51-70), via the pack kernel."""
        _need_cuda(x, "AdaRoundQuantizer.forward")
        n = x.shape[0]
        operand, _, _ = ops.pack_weight(x.detach().float().contiguous(), self.delta, self.zero_point,
                                        self.alpha, float(self.level - 1), True, n_pad=n,
                                        ci_pad=x.shape[1])
        k = operand.shape[1]
        d = self.delta.detach().to(x.device, torch.float32).reshape(n, 1)
        wdq = d * operand.float()
        if x.dim() == 4:  # operand K order is tap-major
            co, ci, kh, kw = x.shape
            wdq = wdq.view(co, kh, kw, ci).permute(0, 3, 1, 2).contiguous()
        return wdq.view_as(x).to(x.dtype)

    def half(self):
        return self

    def float(self):
        return self
