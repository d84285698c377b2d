"""get_qmodel -- same signature and result as the reference's quant/load_qmodel_util.py:28-72."""
from __future__ import annotations

import torch

from .calibration import load_cali_model
from .quant_block import QuantBasicTransformerBlock
from .quant_model import QMODE, QuantModel


def setup_pipe_to_calibrate(model_type, pipe):
    pipe.unet.float()


def setup_pipe_to_inference(model_type, qnn):
    pass


def get_qmodel(model_type, pipe, ckpt_path, wq_params, use_aq, aq_params, softmax_aq_params,
               use_group, num_inference_steps, time_aware_aqtizer):
    if model_type not in ("sd", "sdxl"):
        raise ValueError(f"Unknown model type: {model_type}")
    setup_pipe_to_calibrate(model_type, pipe)
    qnn = QuantModel(model=pipe.unet, wq_params=wq_params, aq_params=aq_params,
                     softmax_aq_params=softmax_aq_params,
                     aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value], tib_recon=False).to("cuda").eval()
    load_cali_model(qnn, init_data=None, use_aq=use_aq, path=ckpt_path,
                    time_aware_aqtizer=time_aware_aqtizer, num_inference_steps=num_inference_steps,
                    use_group=use_group)
    qnn.disable_out_quantization()
    if use_aq:
        for _, module in qnn.named_modules():
            if isinstance(module, QuantBasicTransformerBlock):
                module.attn1.use_aq = True
                module.attn2.use_aq = True
    setup_pipe_to_inference(model_type, qnn)
    return qnn
