"""Synthetic weights and activation scales on the device (bench / smoke set-up; no network means no
pretrained weights or calibration data).  Shapes follow the reference checkpoint schema
(SURVEY.md 8b): K-wise group scales `(1,1,X)` / `(1,X,1)` with `group_num` distinct values for every
3-D quantizer input, scalars for 2-D inputs and `group_num == 1`."""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn as nn


def build_unet(model_type: str, device="cuda", seed: int = 0) -> nn.Module:
    """Random-init UNet2DConditionModel built directly on `device` (uniform +-1/sqrt(fan_in))."""
    from .unet import sd, sdxl
    graph = sdxl if model_type == "sdxl" else sd
    with torch.device("meta"):
        unet = graph.UNet2DConditionModel()
    unet = unet.to_empty(device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for name, p in unet.named_parameters():
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in), generator=g)
            elif name.endswith("weight") and ("norm" in name):
                p.fill_(1.0)
            elif name.endswith("bias") and ("norm" in name):
                p.zero_()
            else:
                p.uniform_(-0.05, 0.05, generator=g)
    return unet


def random_act_tables(qnn, n_steps: int, group_num: int, abits: int, seed: int = 0) -> List[Dict[str, tuple]]:
    """One {quantizer path: (delta, zp)} table per step, SURVEY.md 8d recipe: labels ~ randint(0,g),
    lo = -(U*3+1), hi = U*3+1, delta = (hi-lo)/(L-1), zp = round(-lo/delta); K-wise orientation."""
    from .quant.quant_layer import QuantLayer, UniformAffineQuantizer
    from .unet.common import Attention
    level = 2 ** abits
    g = torch.Generator().manual_seed(seed)
    tables = []
    for _ in range(n_steps):
        tab = {}

        def scales(n, view):
            ng = max(group_num, 1)
            lo = -(torch.rand(ng, generator=g) * 3 + 1)
            hi = torch.rand(ng, generator=g) * 3 + 1
            d = (hi - lo) / (level - 1)
            z = torch.round(-lo / d)
            if n == 0 or group_num <= 1:
                return d[0].clone(), z[0].clone()
            lab = torch.randint(0, ng, (n,), generator=g)
            return d[lab].view(view), z[lab].view(view)

        for path, m in qnn.named_modules():
            if isinstance(m, QuantLayer):
                two_d = any(s in path for s in ("time_embedding", "add_embedding", "time_emb_proj"))
                if two_d:
                    tab[path + ".aqtizer"] = scales(0, None)
                elif m.is_conv:
                    tab[path + ".aqtizer"] = scales(m.w.shape[1] * m.ksize * m.ksize, (1, -1, 1))
                else:
                    tab[path + ".aqtizer"] = scales(m.w.shape[1], (1, 1, -1))
            elif isinstance(m, Attention) and hasattr(m, "aqtizer_q"):
                for qn in ("aqtizer_q", "aqtizer_k", "aqtizer_v"):
                    tab[f"{path}.{qn}"] = scales(m.head_dim, (1, 1, -1))
                if isinstance(m.aqtizer_w, UniformAffineQuantizer):
                    tab[f"{path}.aqtizer_w"] = (torch.tensor(1.0 / (m.aqtizer_w.level - 1)), torch.tensor(0.0))
        tables.append(tab)
    return tables


def make_qmodel(model_type: str, *, wbits: int, abits: int, group_num: int, n_steps: int, log_quant: bool = True,
                real_time: bool = True, start_peak: bool = True, device="cuda", seed: int = 0):
    """Quantized UNet with synthetic weights / scales, assembled through the same calls as
    get_qmodel (QuantModel -> weight-quantizer init -> step tables -> disable_out_quantization)."""
    from .quant.quant_layer import Scaler, QuantLayer, channel_minmax
    from .quant.quant_model import QuantModel, QMODE
    from .quant.quant_block import QuantBasicTransformerBlock
    unet = build_unet(model_type, device, seed)
    qnn = QuantModel(unet, {"bits": wbits, "channel_wise": True, "scaler": Scaler.MINMAX},
                     {"bits": abits, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True},
                     {"softmax_a_bit": abits, "t2i_log_quant": log_quant, "t2i_real_time": real_time,
                      "t2i_start_peak": start_peak, "log_max_1": False},
                     aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value]).to(device).eval()
    qnn.set_quant_state(True, True)
    for m in qnn.modules():
        if isinstance(m, QuantLayer):
            d, z = channel_minmax(m.w, m.wqtizer.level)
            m.wqtizer.delta, m.wqtizer.zero_point, m.wqtizer.init = nn.Parameter(d), nn.Parameter(z), True
    qnn.set_step_tables(random_act_tables(qnn, n_steps, group_num, abits, seed), n_steps)
    qnn.disable_out_quantization()
    for m in qnn.modules():
        if isinstance(m, QuantBasicTransformerBlock):
            m.attn1.use_aq = True
            m.attn2.use_aq = True
    return qnn
