"""load_cali_model -- the checkpoint loader half of the reference's quant/calibration.py
(:208-327).  Calibration itself (cali_model, reconstruction, K-means grouping) stays in the
reference.

Differences from the reference loader, all behaviour-preserving for calibrated checkpoints:
  * no dummy forwards: weight-quantizer (delta, zp) come from the checkpoint (or per-channel MINMAX
    when absent, which is what the dummy forward would compute), activation scales from `act_k`;
  * the checkpoint is read once (the reference calls torch.load three times, :220,:294,:311);
  * time-aware mode uploads every step's scales once (QuantModel.set_step_tables) instead of
    ~750 host->device copies per UNet call (:297-312).
"""
from __future__ import annotations

import logging
from typing import Tuple

import torch
import torch.nn as nn

from .adaptive_rounding import AdaRoundQuantizer, RMODE
from .quant_block import BaseQuantBlock
from .quant_layer import QuantLayer, UniformAffineQuantizer, channel_minmax
from .quant_model import QuantModel

logger = logging.getLogger(__name__)


def uaq2adar(model: nn.Module) -> None:
    """Swap every weight quantizer for an AdaRoundQuantizer (reference :20-43)."""
    for m in model.modules():
        if isinstance(m, QuantLayer) and not m.ignore_recon and not isinstance(m.wqtizer, AdaRoundQuantizer):
            m.wqtizer = AdaRoundQuantizer(m.wqtizer, w=m.original_w.data, rmode=RMODE.LEARNED_HARD_SIGMOID)


def _act_tables(ckpt: dict):
    ks = sorted(int(k[4:]) for k in ckpt if k.startswith("act_"))
    return [ckpt[f"act_{k}"] for k in ks]


def _pairs(act: dict) -> dict:
    """{'model.x.aqtizer': (delta, zp)} from an act_k dict."""
    out = {}
    for key, v in act.items():
        if key.endswith(".delta"):
            path = key[: -len(".delta")]
            out[path] = (v, act[path + ".zero_point"])
    return out


def _is_disabled(named: dict, qpath: str) -> bool:
    """quantizers that never run: conv_in / conv_out (disable_aq) and the log2 map quantizer's uniform twin"""
    owner = named.get(qpath.rpartition(".")[0])
    return bool(getattr(owner, "disable_aq", False))


@torch.no_grad()
def load_cali_model(qnn: QuantModel, init_data: Tuple[torch.Tensor] = None, use_aq: bool = False,
                    path: str = None, time_aware_aqtizer: bool = False, num_inference_steps: int = 25,
                    use_group: bool = False) -> None:
    logger.info("Loading calibration model...")
    full = torch.load(path, map_location="cpu") if isinstance(path, str) else path
    # an un-merged activation file ({'act_0': ..}, reference calibration_group_quantization.py:102-107) carries no
    # weights: the reference then loads nothing into the model (strict=False) and keeps the pipe's own weights
    ckpt = dict(full["weight"]) if "weight" in full else {k: v for k, v in full.items() if not k.startswith("act_")}
    dev = qnn.device

    # weight quantizers: what the first dummy forward initialises (reference :224-225, quant_layer :253-264)
    qnn.set_quant_state(use_wq=True, use_aq=False)
    qnn.disable_out_quantization()
    for m in qnn.model.modules():
        if isinstance(m, QuantLayer):
            d, z = channel_minmax(m.w, m.wqtizer.level)
            m.wqtizer.delta, m.wqtizer.zero_point, m.wqtizer.init = nn.Parameter(d), nn.Parameter(z), True
    if any("alpha" in k for k in ckpt):  # BRECQ checkpoint (reference :227-230)
        uaq2adar(qnn)
        for m in qnn.model.modules():
            if isinstance(m, AdaRoundQuantizer):
                m.delta, m.zero_point = nn.Parameter(m.delta.detach()), nn.Parameter(m.zero_point.detach())
    for key in [k for k in ckpt if "aqtizer" in k]:
        del ckpt[key]
    if ckpt:
        target = qnn if "model" in next(iter(ckpt)) else qnn.model
        missing = target.load_state_dict(ckpt, strict=False)
        logger.info(f"keys not loaded: {missing}")
    qnn.set_quant_state(use_wq=True, use_aq=False)

    if use_aq:
        qnn.set_quant_state(use_wq=True, use_aq=True)
        tables = _act_tables(full)       # act_k are read from the raw file whether or not it is merged (:294, :311)
        if time_aware_aqtizer:
            if not tables:
                raise KeyError("act_0")
            qnn.set_step_tables([_pairs(t) for t in tables], num_inference_steps)
        else:
            act0 = tables[0] if tables else {k: v for k, v in full.items() if "aqtizer" in k}
            named = dict(qnn.named_modules())
            for qpath, (d, z) in _pairs(act0).items():
                qt = named.get(qpath)
                if qt is None:
                    continue
                if d.dim() > 0:
                    if not use_group:  # QDiff branch: load_state_dict onto scalar parameters (reference :314-325)
                        raise RuntimeError(f"size mismatch for {qpath}.delta: checkpoint {tuple(d.shape)} vs ()")
                    owner = named[qpath.rpartition(".")[0]]
                    if isinstance(owner, QuantLayer) and qpath.endswith(".aqtizer"):
                        owner.use_group_num = True           # reference :271-278
                qt.delta = nn.Parameter(d.to(dev))
                if isinstance(qt, UniformAffineQuantizer):
                    qt.zero_point = nn.Parameter(z.to(dev))
                qt.init = True
            # The reference initialises quantizers that the checkpoint does not cover from its RANDOM dummy
            # forward (:255-257) -- values that mean nothing.  Here that is an error, named at load time.
            missing = [p for p, m in named.items()
                       if isinstance(m, UniformAffineQuantizer) and ".aqtizer" in p and m.delta is None
                       and not _is_disabled(named, p)]
            if missing:
                raise KeyError(f"activation checkpoint has no (delta, zero_point) for {len(missing)} quantizers, "
                               f"e.g. {missing[:4]}")
    logger.info("Loading calibration model done.")
