"""QuantModel with the reference's interface (quant/quant_model.py): recursive Conv2d/Linear ->
QuantLayer and Resnet/Transformer -> quant-block surgery, set_quant_state,
disable_out_quantization, the `.config` shim read by the diffusers pipelines.

New here (no counterpart in the reference, which walks ~750 tensors host->device per call,
quant/calibration.py:297-312): `set_step_tables` keeps every step's activation scales resident
on the device and `forward` only flips an index; `capture` records one CUDA graph per step.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .quant_block import (BaseQuantBlock, QuantBasicTransformerBlock, QuantResnetBlock2D, b2qb, T2ILogQuantizer)
from .quant_layer import QMODE, QuantLayer, StraightThrough, UniformAffineQuantizer
from .adaptive_rounding import AdaRoundQuantizer


class CFG:
    in_channels = 0
    sample_size = 0
    time_cond_proj_dim = 0
    addition_time_embed_dim = 0


class QuantModel(nn.Module):
    def __init__(self, model: nn.Module, wq_params: dict = {}, aq_params: dict = {},
                 softmax_aq_params: dict = {}, cali: bool = True, tib_recon: bool = False, **kwargs) -> None:
        super().__init__()
        if tib_recon:
            raise NotImplementedError("tib_recon (TFMQ time-embedding reconstruction) is calibration-only")
        self.model = model
        self.config = CFG()
        self.config.in_channels = model.config.in_channels
        self.config.sample_size = model.config.sample_size
        self.config.time_cond_proj_dim = model.config.time_cond_proj_dim
        if hasattr(model.config, "addition_time_embed_dim"):
            self.config.addition_time_embed_dim = model.config.addition_time_embed_dim
        self.tib_recon = tib_recon
        # constructor arguments and raw activation tables, kept for dgq_b200.compiled.compile_checkpoint
        self._ctor = {"wbits": wq_params.get("bits"), "abits": aq_params.get("bits"),
                      "softmax": {k: softmax_aq_params.get(k) for k in
                                  ("softmax_a_bit", "t2i_log_quant", "t2i_real_time", "t2i_start_peak", "log_max_1")}}
        self._raw_tables = None
        self._device = None
        self.B = b2qb()
        self.quant_module(self.model, wq_params, aq_params,
                          aq_mode=kwargs.get("aq_mode", [QMODE.NORMAL.value]), prev_name=None)
        self.quant_block(self.model, wq_params, aq_params, softmax_aq_params)
        # time-aware state
        self._step_tables: Optional[List[Dict[str, object]]] = None
        self._num_inference_steps = None
        self._graphs = {}
        self._graph_pool = None
        self._use_graphs = False

    # -- tree surgery (reference :66-103) ---------------------------------------------------
    def quant_module(self, module, wq_params={}, aq_params={}, aq_mode=[QMODE.NORMAL.value], prev_name=None):
        for name, child in module.named_children():
            if isinstance(child, tuple(QuantLayer.QMAP.keys())):
                setattr(module, name, QuantLayer(child, wq_params, aq_params, aq_mode=aq_mode))
            elif isinstance(child, StraightThrough):
                continue
            else:
                self.quant_module(child, wq_params, aq_params, aq_mode=aq_mode, prev_name=name)

    def quant_block(self, module, wq_params={}, aq_params={}, softmax_aq_params={}):
        for name, child in module.named_children():
            cls = self.B.get(child.__class__.__name__)
            if cls is QuantBasicTransformerBlock:
                setattr(module, name, cls(child, aq_params, softmax_aq_params))
            elif cls is QuantResnetBlock2D:
                setattr(module, name, cls(child, aq_params))
            else:
                self.quant_block(child, wq_params, aq_params, softmax_aq_params)

    def set_quant_state(self, use_wq: bool = False, use_aq: bool = False) -> None:
        for m in self.model.modules():
            if isinstance(m, (BaseQuantBlock, QuantLayer)):
                m.set_quant_state(use_wq=use_wq, use_aq=use_aq)
        self._graphs.clear()

    def disable_out_quantization(self) -> None:
        self.model.conv_in.use_wq = False
        self.model.conv_in.disable_aq = True
        self.model.conv_out.use_wq = False
        self.model.conv_out.disable_aq = True

    # -- time-aware scales resident on the device -------------------------------------------
    def set_step_tables(self, tables: List[Dict[str, tuple]], num_inference_steps: int) -> None:
        """tables[k] = {module path under self: (delta, zero_point)} for `act_k`.  Every quantizer
        gets one device QParam per step; the sticky use_group_num flip of the reference's loader
        (quant/calibration.py:271-278) is replayed per step, in step order."""
        dev = self.device
        self._raw_tables = tables
        named = dict(self.named_modules())
        qtables: Dict[str, list] = {}
        flips: List[List[str]] = []
        cur_shape: Dict[str, tuple] = {}
        for k, tab in enumerate(tables):
            flip_k = []
            for path, (d, z) in tab.items():
                qt = named.get(path)
                if qt is None or not isinstance(qt, UniformAffineQuantizer):
                    continue
                owner_path, _, leaf = path.rpartition(".")
                owner = named[owner_path]
                conv = leaf == "aqtizer" and isinstance(owner, QuantLayer) and owner.is_conv
                kperm = owner._kperm(dev) if conv else None
                from .. import ops
                qp = ops.qparam_from_ckpt(d, z, float(qt.level - 1), dev, conv=conv, kperm=kperm)
                qtables.setdefault(path, [None] * len(tables))[k] = qp
                if leaf == "aqtizer" and isinstance(owner, QuantLayer):
                    prev = cur_shape.get(path, ())
                    if tuple(d.shape) != prev and owner_path not in [p for f in flips for p in f]:
                        flip_k.append(owner_path)
                    cur_shape[path] = tuple(d.shape)
            flips.append(flip_k)
        for path, lst in qtables.items():
            missing = [k for k, q in enumerate(lst) if q is None]
            if missing:
                raise KeyError(f"{path}: no activation scales for steps {missing}")
            named[path].set_step_table(lst)
        self._step_tables = flips
        self._num_inference_steps = num_inference_steps
        self._named = named
        self._tabled = [named[p] for p in qtables]
        self._cur_step = -1
        self._graphs.clear()

    def step_index(self, timesteps: torch.Tensor) -> int:
        """act_{int((1000 - t) // (1000 // n))} (reference quant/calibration.py:302)."""
        t = timesteps.reshape(-1)[0].item()
        return int((1000 - t) // (1000 // self._num_inference_steps))

    def set_step(self, idx: int) -> None:
        if self._step_tables is None:
            return
        if not 0 <= idx < len(self._step_tables):
            raise KeyError(f"act_{idx}")  # the reference raises KeyError on a missing act_k
        if idx == self._cur_step:
            return
        for k in range(idx + 1):          # sticky flags accumulate in step order
            for owner_path in self._step_tables[k]:
                self._named[owner_path].use_group_num = True
        for m in self._tabled:
            m._step = idx
        self._cur_step = idx

    # -- forward ------------------------------------------------------------------------------
    def forward(self, sample, timesteps, encoder_hidden_states, *args, **kwargs):
        idx = 0
        if self._step_tables is not None:
            idx = self.step_index(timesteps)
            self.set_step(idx)
        if self._use_graphs and sample.is_cuda:
            return self._graph_forward(idx, sample, timesteps, encoder_hidden_states, *args, **kwargs)
        return self.model(sample, timesteps, encoder_hidden_states, *args, **kwargs)

    # -- CUDA graphs: one per (step index, input shapes) --------------------------------------
    def enable_cuda_graphs(self, on: bool = True) -> None:
        """Replay a captured graph per (step, shape) instead of ~3.5k python-dispatched launches.
        Quantizer state must not change afterwards (set_quant_state / set_step_tables clear it)."""
        self._use_graphs = on
        if not on:
            self._graphs.clear()

    def _graph_forward(self, idx, sample, timesteps, ctx, *args, **kwargs):
        added = kwargs.get("added_cond_kwargs", args[0] if args else None)
        key = (idx, tuple(sample.shape), sample.dtype, tuple(ctx.shape), ctx.dtype)
        ent = self._graphs.get(key)
        if ent is None:
            st = {"sample": sample.clone(), "t": timesteps.clone().to(sample.device), "ctx": ctx.clone(),
                  "added": None if added is None else {k: v.clone().to(sample.device) for k, v in added.items()}}

            def call():
                if st["added"] is not None:
                    return self.model(st["sample"], st["t"], st["ctx"], st["added"])
                return self.model(st["sample"], st["t"], st["ctx"])
            from .. import ops
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):     # warm-up: lazy weight packing, allocator pools
                call()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            n0 = ops.LAUNCHES
            # ONE memory pool for every (step, shape) graph of this model: time-aware sampling captures a
            # graph per denoising step (25-50 for SD PLMS) and the graphs replay strictly one after another,
            # so they can share their working set instead of each pinning a private copy of it
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=self._graph_pool):
                out = call()
            ent = {"graph": graph, "st": st, "out": out, "launches": ops.LAUNCHES - n0}
            self._graphs[key] = ent
        st = ent["st"]
        st["sample"].copy_(sample, non_blocking=True)
        st["t"].copy_(timesteps.to(st["t"].dtype), non_blocking=True)
        st["ctx"].copy_(ctx, non_blocking=True)
        if st["added"] is not None:
            for k, v in added.items():
                st["added"][k].copy_(v, non_blocking=True)
        ent["graph"].replay()
        from .. import ops
        ops.LAUNCHES += ent["launches"]
        # the result aliases the graph's static output buffer: it is valid until the next replay of THIS entry
        # (the samplers consume it in their step kernel before the next UNet call); clone it to keep it longer
        return [ent["out"][0]]

    def half(self):
        return self  # compute is fp16 already; parameters and scales stay fp32 (quantisation is done in fp32)

    def float(self):
        return self

    @property
    def device(self):
        if self._device is not None:      # compiled checkpoints keep no master weights (meta parameters)
            return self._device
        return next(self.parameters()).device
