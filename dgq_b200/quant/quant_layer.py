"""QuantLayer / UniformAffineQuantizer with the reference's interface, executed by CUDA kernels.

Mirrors quant/quant_layer.py of ugonfor/DGQ (class and attribute names, constructor signatures,
state-dict keys, set_quant_state semantics) so callers written against the reference keep working;
the arithmetic behind `forward` is dgq_b200's sm_100a kernels:

  UniformAffineQuantizer.forward  -> dgq_fake_quant_f32        (reference :271-299)
  QuantLayer.forward              -> dgq_act_producer / dgq_row_quant + dgq_gemm_f16 (:626-661)
  weight quantisation             -> dgq_pack_weight, once, cached  (reference redoes it per call)

There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import logging
from enum import Enum
from typing import List, Optional, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

logger = logging.getLogger(__name__)


class StraightThrough(nn.Module):
    def forward(self, x):
        return x


def _need_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{what}: dgq_b200 runs on CUDA tensors only (no CPU fallback); got {x.device}")


def minmax(x: torch.Tensor, symmetric: bool = False, level: int = 256, always_zero: bool = False):
    """Scaler.MINMAX (reference quant_layer.py:22-38), asymmetric form only: load-time scale
    initialisation, not on the per-step path."""
    if symmetric:
        raise NotImplementedError("symmetric quantization is not used on DGQ's inference path")
    x_min, x_max = min(float(x.min()), 0.0), max(float(x.max()), 0.0)
    delta = float(x_max - x_min) / (level - 1)
    if always_zero:
        delta = float(x_max) / (level - 1)
    delta = max(delta, 1e-8)
    d = torch.tensor(delta, dtype=torch.float32, device=x.device)
    if always_zero:
        return d, torch.zeros((), dtype=torch.float32, device=x.device)
    zp = torch.round(-torch.tensor(x_min, dtype=torch.float32, device=x.device) / d)
    return d, zp


def _calibration_only(name):
    def f(*a, **k):
        raise NotImplementedError(f"Scaler.{name} belongs to calibration, which stays in the reference "
                                  "(BASELINE.json north_star)")
    f.__name__ = name.lower()
    return f


class Scaler(Enum):
    """Same spelling as the reference: members are plain functions, so `Scaler.MINMAX` is the
    callable itself (reference quant_layer.py:186-192)."""
    MINMAX = minmax
    MSE = _calibration_only("MSE")
    KL = _calibration_only("KL")
    HIST = _calibration_only("HIST")
    OMSE = _calibration_only("OMSE")
    LOGMINMAX = _calibration_only("LOGMINMAX")


QMODE = Enum("QMODE", ("QDIFF", "NORMAL", "PTQD"))


def channel_minmax(w: torch.Tensor, level: int):
    """Vectorised per-out-channel MINMAX (reference :253-264 loops over channels in python)."""
    flat = w.detach().reshape(w.shape[0], -1).double()
    lo = torch.clamp(flat.amin(dim=1), max=0.0)
    hi = torch.clamp(flat.amax(dim=1), min=0.0)
    delta = ((hi - lo) / (level - 1)).float()
    delta = torch.where(delta < 1e-8, torch.full_like(delta, 1e-8), delta)
    zp = torch.round(-lo.float() / delta)
    shape = (-1,) + (1,) * (w.dim() - 1)
    return delta.view(shape), zp.view(shape)


class UniformAffineQuantizer(nn.Module):
    """Asymmetric uniform quantizer; delta / zero_point are (), (1,1,X), (1,X,1) or per-out-channel."""

    def __init__(self, bits: int = 8, symmetric: bool = False, channel_wise: bool = False,
                 scaler=Scaler.MINMAX, leaf_param: bool = False, always_zero: bool = False,
                 quant_emb: bool = False) -> None:
        super().__init__()
        if symmetric:
            raise NotImplementedError("symmetric quantization is not used on DGQ's inference path")
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
        self.group_num = -1
        # time-aware tables: one device-resident QParam per denoising step (replaces the per-call
        # host walk + ~750 H2D copies of reference calibration.py:297-312)
        self._table = None
        self._step = 0
        self._qcache = {}

    # -- parameter plumbing -----------------------------------------------------------------
    def _init_quantization_param(self, x: torch.Tensor, channel_wise: bool = False):
        if channel_wise:
            return channel_minmax(x, self.level)
        return self.scaler(x, self.symmetric, self.level, self.always_zero)

    def _ensure_init(self, x: torch.Tensor) -> None:
        if not self.init or self.delta is None:
            self.delta, self.zero_point = self._init_quantization_param(x, self.channel_wise)
            if self.leaf_param:
                self.delta = nn.Parameter(self.delta)
            self.init = True

    def set_step_table(self, table) -> None:
        self._table = table

    def qparam(self, device, *, conv: bool = False, kperm=None) -> ops.QParam:
        """Device-resident (delta, zp) for the fused kernels; cached until the tensors change."""
        if self._table is not None:
            return self._table[self._step]
        d, z = self.delta, self.zero_point
        if d is None:
            raise RuntimeError("quantizer used before its (delta, zero_point) were loaded or initialised")
        if not torch.is_tensor(z):
            z = torch.tensor(float(z))
        key = (id(d), d._version, id(z), z._version, conv, str(device))
        hit = self._qcache.get("k")
        if hit is None or hit[0] != key:
            hit = (key, ops.qparam_from_ckpt(d, z, float(self.level - 1), device, conv=conv, kperm=kperm))
            self._qcache["k"] = hit
        return hit[1]

    # -- stand-alone forward (reference :271-299) -------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _need_cuda(x, "UniformAffineQuantizer.forward")
        self._ensure_init(x)
        d = self.delta.detach().to(device=x.device, dtype=torch.float32)
        z = self.zero_point
        z = (z.detach() if torch.is_tensor(z) else torch.tensor(float(z))).to(device=x.device, dtype=torch.float32)
        xf = x.detach().to(torch.float32).contiguous()
        # which axis of x the parameter tensor follows (the reference relies on broadcasting)
        period, inner = 1, 1
        if d.numel() > 1:
            shp = list(d.shape)
            while len(shp) < xf.dim():
                shp.insert(0, 1)
            axes = [i for i, s in enumerate(shp) if s != 1]
            if len(axes) != 1 or shp[axes[0]] != xf.shape[axes[0]]:
                raise ValueError(f"cannot broadcast quantizer parameters {tuple(d.shape)} over {tuple(x.shape)}")
            period = xf.shape[axes[0]]
            inner = 1
            for s in xf.shape[axes[0] + 1:]:
                inner *= s
        dd = d.reshape(-1).contiguous()
        zz = z.reshape(-1).expand(dd.numel()).contiguous()
        out = ops.fake_quant(xf, dd, zz, period, inner, float(self.level - 1))
        return out.to(x.dtype)

    def codes(self, x: torch.Tensor) -> torch.Tensor:
        """Integer codes (uint8) of `x` -- verification helper, same kernel as forward."""
        _need_cuda(x, "UniformAffineQuantizer.codes")
        d = self.delta.detach().to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
        z = self.zero_point
        z = (z.detach() if torch.is_tensor(z) else torch.tensor(float(z))).to(x.device, torch.float32)
        if d.numel() != 1:
            raise NotImplementedError
        _, c = ops.fake_quant(x.detach().float().contiguous(), d, z.reshape(1), 1, 1, float(self.level - 1),
                              want_codes=True)
        return c

    def bitwidth_refactor(self, bits: int = 8) -> None:
        self.level = 2 ** bits

    def extra_repr(self) -> str:
        return (f"level={self.level}, symmetric={self.symmetric}, channel_wise={self.channel_wise}, "
                f"leaf_param={self.leaf_param}, group_num={self.group_num}")

    def half(self):
        return self  # parameters stay fp32: the kernels quantise in fp32 and compute in fp16

    def float(self):
        return self


class QuantLayer(nn.Module):
    """Drop-in for the reference's QuantLayer (quant/quant_layer.py:577-702)."""

    QMAP = {nn.Conv2d: F.conv2d, nn.Linear: F.linear}

    def __init__(self, layer: Union[nn.Conv2d, nn.Linear], wq_params: dict = {}, aq_params: dict = {},
                 disable_aq: bool = False, aq_mode: List[int] = [QMODE.QDIFF.value],
                 quant_emb: bool = False) -> None:
        super().__init__()
        self.wq_params = dict(wq_params)
        self.aq_params = dict(aq_params)
        self.fwd_kwargs = {}
        if isinstance(layer, nn.Conv2d):
            if layer.groups != 1 or layer.dilation != (1, 1) or layer.kernel_size[0] != layer.kernel_size[1]:
                raise NotImplementedError("only square, dense, undilated convolutions occur in the DGQ UNets")
            if layer.kernel_size[0] not in (1, 3) or layer.padding[0] != layer.kernel_size[0] // 2:
                raise NotImplementedError("conv kernels are 1x1 (pad 0) or 3x3 (pad 1)")
            self.fwd_kwargs = dict(stride=layer.stride, padding=layer.padding, dilation=layer.dilation,
                                   groups=layer.groups)
        self.kwd_func = self.QMAP[type(layer)]
        self.w = layer.weight
        self.original_w = self.w.data.clone()
        self.b = None
        self.original_b = None
        if layer.bias is not None:
            self.b = layer.bias
            self.original_b = self.b.data.clone()
        self.use_wq = False
        self.use_aq = False
        self.disable_aq = disable_aq
        self.aq_mode = aq_mode
        self.quant_emb = quant_emb
        self.wq_params["quant_emb"] = quant_emb
        self.wqtizer = UniformAffineQuantizer(**self.wq_params)
        self.aqtizer = UniformAffineQuantizer(**self.aq_params)
        self.split = 0
        self.act_func = StraightThrough()
        self.ignore_recon = False
        self.extra_repr = layer.extra_repr
        self.use_group_num = False
        self._pack = None
        self._frozen = None   # (operand, scale, bias, n_pad) installed by dgq_b200.compiled.load_compiled

    # -- geometry -----------------------------------------------------------------------------
    @property
    def is_conv(self) -> bool:
        return self.kwd_func is F.conv2d

    @property
    def ksize(self) -> int:
        return self.w.shape[2] if self.is_conv else 1

    @property
    def stride(self) -> int:
        return self.fwd_kwargs["stride"][0] if self.is_conv else 1

    @property
    def out_features(self) -> int:
        return self.w.shape[0]

    def _kperm(self, device):
        """reference unfold order (c*k*k + tap) -> GEMM K order (tap*C_pad + c)."""
        if not self.is_conv or self.ksize == 1:
            return None
        hit = self.__dict__.get("_kperm_cache")
        if hit is None or hit[0] != str(device):   # built once: three tiny kernels per conv per call otherwise
            ci, kk = self.w.shape[1], self.ksize * self.ksize
            hit = (str(device), (torch.arange(ci, device=device).view(1, ci) * kk
                                 + torch.arange(kk, device=device).view(kk, 1)).reshape(-1))
            self.__dict__["_kperm_cache"] = hit
        return hit[1]

    # -- packed weights (K2: once, not per forward) -----------------------------------------
    def i8_ok(self, q: "ops.QParam") -> bool:
        """this layer can run on the kind::i8 qGEMM under activation quantizer q: quantized W4 / W8 weights, an
        activation scale that is constant along K (scalar / row-wise), byte operands with 16-byte rows, and -- for
        a conv on the per-tensor path, whose zero padding is an exact 0 (reference :659) -- a zero point that is
        itself a code (0 <= zp <= qmax), so that padding taps can be written as the code zp."""
        if not (self.use_wq and q.exact and self.wqtizer.level in (16, 256)) or self._frozen is not None:
            return False
        k = self.w[0].numel()
        if k % 16:
            return False
        if self.is_conv and self.ksize > 1 and not self.pad_quantized and not q.zp_in_range:
            return False
        return True

    def packed(self, geglu: bool = False, i8: bool = False):
        """(operand fp16 [n_pad, K], scale fp32 [n_pad] | None, bias fp32 [n_pad] | None, n_pad).
        geglu: rows interleaved for the fused GEGLU epilogue (cached separately).
        i8: the s8 operand of dgq_gemm_i8 instead, + (colsum int32 [n_pad], b_off int32 [n_pad] | None,
        csoob int32 [9, n_pad] | None: border-class tables of the implicit 3x3 conv)."""
        if i8:
            return self._packed_i8(geglu)
        if self._frozen is not None:
            return self._frozen_pack(geglu)
        w = self.w if self.use_wq else self.original_w
        b = self.b if self.use_wq else self.original_b
        _need_cuda(w, "QuantLayer weights")
        wq = self.wqtizer
        alpha = getattr(wq, "alpha", None)
        key = (self.use_wq, id(w), w._version, None if b is None else (id(b), b._version), geglu)
        if self.use_wq:
            if wq.delta is None:
                wq.delta, wq.zero_point = channel_minmax(self.w, wq.level)  # reference :253-264
                wq.init = True
            key += (id(wq.delta), wq.delta._version, id(wq.zero_point), wq.zero_point._version,
                    None if alpha is None else (id(alpha), alpha._version))
        hit = self._pack.get(geglu) if self._pack else None
        if hit is not None and hit[0] == key:
            return hit[1]
        dev = w.device
        n = w.shape[0]
        n_pad = (n + 7) // 8 * 8
        w32 = w.detach().to(torch.float32)
        operand, _, _ = ops.pack_weight(w32, wq.delta if self.use_wq else None,
                                        wq.zero_point if self.use_wq else None,
                                        alpha if self.use_wq else None, float(wq.level - 1), self.use_wq,
                                        n_pad=n_pad)
        scale = None
        if self.use_wq:
            scale = torch.zeros(n_pad, dtype=torch.float32, device=dev)
            scale[:n] = wq.delta.detach().reshape(-1).to(dev, torch.float32)
        bias = None
        if b is not None:
            bias = torch.zeros(n_pad, dtype=torch.float32, device=dev)
            bias[:n] = b.detach().to(dev, torch.float32)
        if geglu:
            # GEGLU projection feeding the fused epilogue (DGQ_EPI_GEGLU): GEMM columns are re-ordered
            # [32 x1 | 32 gate] per 64, so one epilogue thread holds x1 and its gate together
            f = n // 2
            if n % 64 or n_pad != n:
                raise ValueError("GEGLU interleave needs 2f to be a multiple of 64")
            i = torch.arange(n, device=dev)
            perm = (i // 64) * 32 + i % 32 + ((i % 64) >= 32) * f
            operand = operand[perm].contiguous()
            scale = None if scale is None else scale[perm].contiguous()
            bias = None if bias is None else bias[perm].contiguous()
        if not self._pack:
            self._pack = {}
        self._pack[geglu] = (key, (operand, scale, bias, n_pad))
        return self._pack[geglu][1]

    def _packed_i8(self, geglu: bool):
        """kind::i8 operand: s8 (code - off_n), off_n = zp_n for W4 (exact, nothing left to correct) or 128 for W8,
        with the per-channel column sums and residual offsets of dgq_gemm_i8's integer epilogue."""
        w, b, wq = self.w, self.b, self.wqtizer
        _need_cuda(w, "QuantLayer weights")
        if wq.delta is None:
            wq.delta, wq.zero_point = channel_minmax(self.w, wq.level)
            wq.init = True
        alpha = getattr(wq, "alpha", None)
        key = ("i8", id(w), w._version, None if b is None else (id(b), b._version), geglu, id(wq.delta),
               wq.delta._version, id(wq.zero_point), wq.zero_point._version,
               None if alpha is None else (id(alpha), alpha._version))
        hit = self._pack.get(("i8", geglu)) if self._pack else None
        if hit is not None and hit[0] == key:
            return hit[1]
        dev, n = w.device, w.shape[0]
        n_pad = (n + 7) // 8 * 8
        qmax = float(wq.level - 1)
        _, codes, _ = ops.pack_weight(w.detach().to(torch.float32), wq.delta, wq.zero_point, alpha, qmax, True,
                                      n_pad=n_pad, want_codes=True)
        operand, colsum, b_off = ops.weight_to_i8(codes, wq.zero_point, n, qmax)
        # implicit-conv border tables (3x3 only; GEGLU never applies to a conv)
        csoob = ops.conv_oob_colsum(operand, w.shape[1]) if (self.is_conv and self.ksize == 3) else None
        scale = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        scale[:n] = wq.delta.detach().reshape(-1).to(dev, torch.float32)
        bias = None
        if b is not None:
            bias = torch.zeros(n_pad, dtype=torch.float32, device=dev)
            bias[:n] = b.detach().to(dev, torch.float32)
        if geglu:
            if n % 64 or n_pad != n:
                raise ValueError("GEGLU interleave needs 2f to be a multiple of 64")
            i = torch.arange(n, device=dev)
            perm = (i // 64) * 32 + i % 32 + ((i % 64) >= 32) * (n // 2)
            operand, scale, colsum = operand[perm].contiguous(), scale[perm].contiguous(), colsum[perm].contiguous()
            bias = None if bias is None else bias[perm].contiguous()
            b_off = None if b_off is None else b_off[perm].contiguous()
        if not self._pack:
            self._pack = {}
        self._pack[("i8", geglu)] = (key, (operand, scale, bias, n_pad, colsum, b_off, csoob))
        return self._pack[("i8", geglu)][1]

    def _frozen_pack(self, geglu: bool):
        """Operands of a compiled checkpoint (no fp32 master weights on the device)."""
        if not geglu:
            return self._frozen
        hit = self._pack.get("frozen_geglu") if self._pack else None
        if hit is None:
            operand, scale, bias, n_pad = self._frozen
            n = self.out_features
            if n % 64 or n_pad != n:
                raise ValueError("GEGLU interleave needs 2f to be a multiple of 64")
            i = torch.arange(n, device=operand.device)
            perm = (i // 64) * 32 + i % 32 + ((i % 64) >= 32) * (n // 2)
            hit = (operand[perm].contiguous(), None if scale is None else scale[perm].contiguous(),
                   None if bias is None else bias[perm].contiguous(), n_pad)
            self._pack = dict(self._pack or {}, frozen_geglu=hit)
        return hit

    def packed_int4(self):
        """The W4 checkpoint payload: two codes per byte + per-channel (delta, zp) -- 0.5 B/weight."""
        wq = self.wqtizer
        if wq.level != 16:
            raise ValueError("packed_int4 needs a 4-bit weight quantizer")
        _, _, packed = ops.pack_weight(self.w.detach().float(), wq.delta, wq.zero_point, getattr(wq, "alpha", None),
                                       15.0, True, want_packed4=True)
        return packed

    # -- activation quantizer of this layer -----------------------------------------------
    def act_qparam(self, device) -> ops.QParam:
        if not (self.use_aq and not self.disable_aq):
            return ops.NOQ
        if self.aqtizer._table is None and self.aqtizer.delta is None:
            raise RuntimeError("activation quantizer has no parameters: load a calibration checkpoint "
                               "(quant.load_qmodel_util.get_qmodel) or run the stand-alone forward once")
        if self.aqtizer._table is not None:        # time-aware: device tables already in GEMM K order
            return self.aqtizer._table[self.aqtizer._step]
        return self.aqtizer.qparam(device, conv=self.is_conv, kperm=self._kperm(device))

    @property
    def pad_quantized(self) -> bool:
        """unfold path quantises the zero padding too (reference :630-641)."""
        return bool(self.use_group_num and self.is_conv)

    # -- stand-alone forward (reference :626-661) -------------------------------------------
    def forward(self, x: torch.Tensor, split: int = 0) -> torch.Tensor:
        from .. import engine
        _need_cuda(x, "QuantLayer.forward")
        if self.use_aq and not self.disable_aq and self.aqtizer._table is None and self.aqtizer.delta is None:
            xin = x
            if self.use_group_num and self.is_conv:
                xin = F.unfold(x, self.ksize, padding=self.ksize // 2, stride=self.stride)
            self.aqtizer._ensure_init(xin)  # reference :274-278: first forward initialises from data
        return engine.quant_layer_forward(self, x)

    def set_quant_state(self, use_wq: bool = False, use_aq: bool = False) -> None:
        self.use_wq = use_wq if not self.ignore_recon else False
        self.use_aq = use_aq if not self.ignore_recon else False

    def set_running_stat(self, running_stat: bool) -> None:
        self.aqtizer.running_stat = running_stat

    def _calibration(self, *a, **k):
        raise NotImplementedError("group calibration stays in the reference (BASELINE.json north_star)")

    set_group_num = _calibration
    done_group_num = _calibration

    def half(self):
        return self

    def float(self):
        return self

    def _apply(self, fn, *args, **kwargs):
        # keep the plain-tensor copies (not registered as buffers in the reference either) on the
        # module's device, as the reference does by hand in forward (`w.to(x.device)`, :648-650)
        super()._apply(fn, *args, **kwargs)
        self.original_w = fn(self.original_w)
        if self.original_b is not None:
            self.original_b = fn(self.original_b)
        self._pack = None
        return self
