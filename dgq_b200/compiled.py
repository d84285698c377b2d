"""Compiled checkpoints: one file holding what the engine keeps resident.

The reference start-up is `torch.load` of multi-GB fp32 pickles (three times), a module-tree rewrite, two
dummy forwards and a re-quantisation of every weight on every call (quant/calibration.py:208-327,
quant/quant_layer.py:642-643).  `compile_checkpoint` turns a loaded QuantModel into

    DGQB2001 | header length | JSON header | 64-byte aligned raw tensors

with, per QuantLayer, the packed integer weight codes (two 4-bit codes per byte for W4, one byte for W8, in
the GEMM's K order), the per-out-channel (delta, zero_point), the bias; fp16 weights for the layers that stay
FP (conv_in / conv_out); the norm parameters; and every step's activation-quantizer (delta, zero_point) in
the reference checkpoint's own shapes.  The header carries the model type, bit widths, softmax-quantizer
switches, per-layer flags and a SHA-256 of the payload.

`load_compiled` builds the module tree on the meta device (no random init, no fp32 master weights), unpacks
the codes on the GPU (dgq_unpack_weight) and installs the step tables -- the result computes bit-identical
outputs to the QuantModel it was compiled from.  The model is inference-only: `state_dict()` holds meta
tensors for the weights.
"""
from __future__ import annotations

import hashlib
import json
import struct
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn as nn

MAGIC = b"DGQB2001"
_DT = {"u8": np.uint8, "f16": np.float16, "f32": np.float32}


def _np(t: torch.Tensor) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy())


def compile_checkpoint(qnn, path: str) -> dict:
    """Write `qnn` (a CUDA QuantModel with weights, weight-quantizer parameters and activation scales loaded,
    e.g. the result of quant.load_qmodel_util.get_qmodel) to `path`.  Returns the header."""
    from . import ops
    from .quant.quant_layer import QuantLayer, UniformAffineQuantizer
    from .unet.common import Attention
    tensors: Dict[str, np.ndarray] = {}
    layers: Dict[str, dict] = {}
    attn: Dict[str, dict] = {}
    norms: List[str] = []
    for name, m in qnn.named_modules():
        if isinstance(m, QuantLayer):
            wq = m.wqtizer
            w = m.w if m.use_wq else m.original_w
            b = m.b if m.use_wq else m.original_b
            if not w.is_cuda:
                raise RuntimeError("compile_checkpoint needs the model on a CUDA device")
            n = w.shape[0]
            n_pad = (n + 7) // 8 * 8
            ci = w.shape[1]
            ci_pad = (ci + 7) // 8 * 8
            taps = m.ksize * m.ksize
            ent = {"use_wq": bool(m.use_wq), "use_aq": bool(m.use_aq), "disable_aq": bool(m.disable_aq),
                   "use_group_num": bool(m.use_group_num), "n": n, "n_pad": n_pad, "ci": ci, "ci_pad": ci_pad, "taps": taps,
                   "bias": b is not None}
            if m.use_wq:
                if wq.delta is None:
                    raise RuntimeError(f"{name}: weight quantizer has no (delta, zero_point)")
                bits = 4 if wq.level == 16 else 8
                if wq.level > 256:
                    raise ValueError("weight codes wider than 8 bits")
                _, codes, packed4 = ops.pack_weight(w.detach().float(), wq.delta, wq.zero_point, getattr(wq, "alpha", None),
                                                    float(wq.level - 1), True, n_pad=n_pad, want_codes=True,
                                                    want_packed4=bits == 4)
                ent["bits"] = bits
                tensors[name + ".codes"] = _np(packed4 if bits == 4 else codes)
                tensors[name + ".wdelta"] = _np(wq.delta.reshape(-1).float())
                tensors[name + ".wzp"] = _np(wq.zero_point.reshape(-1).float().expand(n))
            else:
                operand, _, _ = ops.pack_weight(w.detach().float(), None, None, None, 0.0, False, n_pad=n_pad)
                ent["bits"] = 16
                tensors[name + ".w16"] = _np(operand)
            if b is not None:
                tensors[name + ".b"] = _np(b.float())
            layers[name] = ent
        elif isinstance(m, (nn.GroupNorm, nn.LayerNorm)):
            norms.append(name)
            tensors[name + ".weight"] = _np(m.weight.float())
            tensors[name + ".bias"] = _np(m.bias.float())
        elif isinstance(m, Attention) and hasattr(m, "aqtizer_q"):
            attn[name] = {"use_aq": bool(getattr(m, "use_aq", False))}
    # activation scales: per-step tables (time-aware) or the quantizers' own parameters as one table
    if qnn._raw_tables is not None:
        tables, time_aware = qnn._raw_tables, True
    else:
        tab = {}
        for name, m in qnn.named_modules():
            if isinstance(m, UniformAffineQuantizer) and not name.endswith("wqtizer") and m.delta is not None:
                tab[name] = (m.delta, m.zero_point if torch.is_tensor(m.zero_point) else torch.tensor(float(m.zero_point)))
            elif name.endswith("aqtizer_w") and getattr(m, "delta", None) is not None and not isinstance(m, UniformAffineQuantizer):
                tab[name] = (m.delta, torch.zeros(()))
        tables, time_aware = [tab], False
    act_index: List[Dict[str, list]] = []
    for k, tab in enumerate(tables):
        idx = {}
        for qpath, (d, z) in tab.items():
            tensors[f"act.{k}.{qpath}.delta"] = _np(torch.as_tensor(d).float())
            tensors[f"act.{k}.{qpath}.zero_point"] = _np(torch.as_tensor(z).float())
            idx[qpath] = list(torch.as_tensor(d).shape)
        act_index.append(idx)

    index, off, h = {}, 0, hashlib.sha256()
    blobs = []
    for key, arr in tensors.items():
        dt = {np.dtype(np.uint8): "u8", np.dtype(np.float16): "f16", np.dtype(np.float32): "f32"}[arr.dtype]
        raw = arr.tobytes()
        pad = (-len(raw)) % 64
        index[key] = {"dtype": dt, "shape": list(arr.shape), "offset": off, "nbytes": len(raw)}
        blobs.append(raw + b"\0" * pad)
        h.update(raw)
        off += len(raw) + pad
    header = {"format": 1, "model_type": "sdxl" if hasattr(qnn.model, "add_embedding") else "sd", **qnn._ctor,
              "time_aware": time_aware, "num_inference_steps": qnn._num_inference_steps, "layers": layers, "attention": attn,
              "norms": norms, "act": act_index, "tensors": index, "payload_bytes": off, "sha256": h.hexdigest()}
    hj = json.dumps(header).encode()
    hj += b" " * ((-(len(MAGIC) + 8 + len(hj))) % 64)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for bl in blobs:
            f.write(bl)
    return header


def read_header(path: str) -> Tuple[dict, int]:
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError(f"{path}: not a dgq_b200 compiled checkpoint")
        (n,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(n).decode())
    if header.get("format") != 1:
        raise ValueError(f"{path}: unsupported format {header.get('format')}")
    return header, len(MAGIC) + 8 + n


def load_compiled(path: str, device="cuda", verify: bool = True):
    """QuantModel (eval, on `device`) from a compiled checkpoint; raises ValueError on a corrupt payload."""
    from . import ops
    from .quant.quant_layer import QuantLayer, Scaler
    from .quant.quant_model import QMODE, QuantModel
    from .unet import sd, sdxl
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("dgq_b200 runs on CUDA devices only (no CPU fallback)")
    header, base = read_header(path)
    payload = np.memmap(path, dtype=np.uint8, mode="r", offset=base, shape=(header["payload_bytes"],))
    if verify:
        h = hashlib.sha256()
        for ent in header["tensors"].values():
            h.update(payload[ent["offset"]: ent["offset"] + ent["nbytes"]].tobytes())
        if h.hexdigest() != header["sha256"]:
            raise ValueError(f"{path}: payload hash mismatch (corrupt or truncated file)")

    def get(key: str) -> torch.Tensor:
        ent = header["tensors"][key]
        arr = np.frombuffer(payload, dtype=_DT[ent["dtype"]], count=int(np.prod(ent["shape"], dtype=np.int64)),
                            offset=ent["offset"]).reshape(ent["shape"])
        return torch.from_numpy(np.array(arr))          # private copy: the memmap goes away

    graph = sdxl if header["model_type"] == "sdxl" else sd
    sm = header["softmax"]
    with torch.device("meta"):
        unet = graph.UNet2DConditionModel()
        qnn = QuantModel(unet, {"bits": header["wbits"], "channel_wise": True, "scaler": Scaler.MINMAX},
                         {"bits": header["abits"], "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True},
                         dict(sm), aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    qnn._device = device
    named = dict(qnn.named_modules())
    for name in header["norms"]:
        m = named[name]
        m.weight = nn.Parameter(get(name + ".weight").to(device), requires_grad=False)
        m.bias = nn.Parameter(get(name + ".bias").to(device), requires_grad=False)
    for name, ent in header["layers"].items():
        m = named[name]
        if not isinstance(m, QuantLayer):
            raise ValueError(f"{name}: checkpoint layer does not exist in the {header['model_type']} graph")
        n, n_pad = ent["n"], ent["n_pad"]
        if (m.out_features, m.w.shape[1], m.ksize * m.ksize) != (n, ent["ci"], ent["taps"]):
            raise ValueError(f"{name}: shape mismatch with the {header['model_type']} graph")
        scale = None
        if ent["bits"] == 16:
            operand = get(name + ".w16").to(device)
        else:
            operand = ops.unpack_weight(get(name + ".codes").to(device), ent["bits"], get(name + ".wzp").to(device), n,
                                        ent["ci"], ent["taps"], ent["ci_pad"], n_pad)
            scale = torch.zeros(n_pad, dtype=torch.float32, device=device)
            scale[:n] = get(name + ".wdelta").to(device)
        bias = None
        if ent["bias"]:
            bias = torch.zeros(n_pad, dtype=torch.float32, device=device)
            bias[:n] = get(name + ".b").to(device)
        m._frozen = (operand, scale, bias, n_pad)
        m.use_wq, m.use_aq, m.disable_aq = ent["use_wq"], ent["use_aq"], ent["disable_aq"]
        m.use_group_num = ent["use_group_num"]
    for name, ent in header["attention"].items():
        named[name].use_aq = ent["use_aq"]
    tables = [{qp: (get(f"act.{k}.{qp}.delta"), get(f"act.{k}.{qp}.zero_point")) for qp in idx}
              for k, idx in enumerate(header["act"])]
    if header["time_aware"]:
        qnn.set_step_tables(tables, header["num_inference_steps"])
    else:
        for qp, (d, z) in tables[0].items():
            qt = named[qp]
            qt.delta = nn.Parameter(d.to(device), requires_grad=False)
            if hasattr(qt, "zero_point"):
                qt.zero_point = nn.Parameter(z.to(device), requires_grad=False)
            qt.init = True
    return qnn.eval()
