"""Shared construction of the full-UNet parity cases (same recipe as tests/golden/make_golden.py,
which ran the reference on them and stored its output latents)."""
import os
import sys
import types

import torch

from oracle import dgq_oracle as O, synth as S
from tests.golden.make_golden import UNET_CASES, UNET_RUNS, build_case  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(model_type, case):
    return torch.load(os.path.join(GOLDEN, f"unet_{model_type}_{case}.pt"))


def case_inputs(model_type, case, k):
    wb, ab, gn, log, rt, sp, n_steps, batch, ts = UNET_CASES[case]
    return S.example_inputs(model_type, batch, seed=k, t=ts[k])


def oracle_outputs(model_type, case):
    """Run the CPU oracle on every step of a case; returns (outs, sd, cfg, acts)."""
    sd, cfg, acts = build_case(S, O, model_type, case, torch)
    n_steps = UNET_CASES[case][6]
    outs = []
    for k in range(n_steps):
        O.update_group_convs(cfg, acts[k], sd)      # sticky use_group_num, in step order
        inp = case_inputs(model_type, case, k)
        with torch.no_grad():
            outs.append(O.unet_forward(model_type, sd, acts[k], cfg, *inp))
    return outs, sd, cfg, acts


def plain_state(sd):
    """QuantModel-schema keys -> plain UNet keys (.w/.b -> .weight/.bias), wqtizer entries dropped."""
    plain = {}
    for k, v in sd.items():
        if "wqtizer" in k:
            continue
        k2 = k[len("model."):]
        if k2.endswith(".w"):
            k2 = k2[:-2] + ".weight"
        elif k2.endswith(".b"):
            k2 = k2[:-2] + ".bias"
        plain[k2] = v
    return plain


def build_qmodel(model_type, case, sd, acts, tmp_path, device="cuda"):
    """dgq_b200's drop-in path: UNet2DConditionModel() -> load weights -> get_qmodel(ckpt)."""
    from quant.quant_layer import Scaler
    from quant.load_qmodel_util import get_qmodel
    from dgq_b200.unet import sd as sd_graph, sdxl as sdxl_graph
    wb, ab, gn, log, rt, sp, n_steps, batch, ts = UNET_CASES[case]
    ckpt = {"weight": sd}
    for k, a in enumerate(acts):
        ckpt[f"act_{k}"] = a
    path = os.path.join(str(tmp_path), f"{model_type}_{case}.pth")
    torch.save(ckpt, path)
    graph = sdxl_graph if model_type == "sdxl" else sd_graph
    with torch.device("meta"):
        unet = graph.UNet2DConditionModel()
    unet = unet.to_empty(device=device)
    unet.load_state_dict({k: v.to(device) for k, v in plain_state(sd).items()}, strict=True)
    pipe = types.SimpleNamespace(unet=unet)
    qnn = get_qmodel(model_type, pipe, path,
                     {"bits": wb, "channel_wise": True, "scaler": Scaler.MINMAX}, True,
                     {"bits": ab, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True},
                     {"softmax_a_bit": ab, "t2i_log_quant": log, "t2i_real_time": rt, "t2i_start_peak": sp,
                      "log_max_1": False},
                     gn > 1, n_steps, True)
    return qnn


def run_qmodel(qnn, model_type, case, k, device="cuda"):
    inp = case_inputs(model_type, case, k)
    with torch.no_grad():
        if model_type == "sdxl":
            added = {kk: v.to(device) for kk, v in inp[3].items()}
            return qnn(inp[0].to(device), inp[1].to(device), inp[2].to(device), added_cond_kwargs=added)[0]
        return qnn(inp[0].to(device), inp[1].to(device), inp[2].to(device))[0]


def cosine(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten().float().cpu(), b.flatten().float().cpu(), dim=0).item()
