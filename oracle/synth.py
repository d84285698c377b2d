"""Synthetic checkpoints in the reference's exact ``.pth`` schema -- TEST INFRASTRUCTURE.

No pretrained weights or calibration data are reachable (no network), so the
parity tests, ``smoke()`` and ``bench.py`` mint their own inputs here:

* ``make_weights``     -- random-init UNet weights keyed like the reference's
                          ``QuantModel.state_dict()``, values drawn
                          from a numpy PCG64 stream seeded per tensor name, so the
                          same bytes come out on every machine;
* ``init_weight_quant``-- the per-out-channel MINMAX (delta, zp) the reference's
                          loader computes in its first dummy forward
                          (quant/quant_layer.py:253-264, quant/calibration.py:224-225);
* ``calibrate_act``    -- distribution-aware group scales from recorded min/max,
                          following done_group_num (quant/quant_layer.py:315-429)
                          with a deterministic range-sort clustering standing in for
                          K-means (calibration itself is out of scope);
* ``random_act``       -- SURVEY.md 8d's cheap synthetic scales (random labels and
                          ranges), used by bench.py where values do not matter.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import dgq_oracle as O

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# parameter enumeration (names and shapes of QuantModel.state_dict())
# --------------------------------------------------------------------------- #
def _resnet(p, cin, cout, shortcut):
    yield p + ".norm1", ("gn", cin)
    yield p + ".conv1", ("conv", cout, cin, 3)
    yield p + ".time_emb_proj", ("lin", cout, 1280, True)
    yield p + ".norm2", ("gn", cout)
    yield p + ".conv2", ("conv", cout, cout, 3)
    if shortcut:
        yield p + ".conv_shortcut", ("conv", cout, cin, 1)


def _attn(p, c, ctx_dim):
    yield p + ".to_q", ("lin", c, c, False)
    yield p + ".to_k", ("lin", c, ctx_dim, False)
    yield p + ".to_v", ("lin", c, ctx_dim, False)
    yield p + ".to_out.0", ("lin", c, c, True)


def _tblock(p, c, ctx_dim):
    yield p + ".norm1", ("ln", c)
    yield from _attn(p + ".attn1", c, c)
    yield p + ".norm2", ("ln", c)
    yield from _attn(p + ".attn2", c, ctx_dim)
    yield p + ".norm3", ("ln", c)
    yield p + ".ff.net.0.proj", ("lin", 8 * c, c, True)
    yield p + ".ff.net.2", ("lin", c, 4 * c, True)


def _t2d(p, c, n_layers, ctx_dim, linear_proj):
    yield p + ".norm", ("gn", c)
    yield p + ".proj_in", (("lin", c, c, True) if linear_proj else ("conv", c, c, 1))
    for i in range(n_layers):
        yield from _tblock(f"{p}.transformer_blocks.{i}", c, ctx_dim)
    yield p + ".proj_out", (("lin", c, c, True) if linear_proj else ("conv", c, c, 1))


def iter_modules(model_type: str) -> Iterator[Tuple[str, tuple]]:
    """Yield (module path under ``model.``, descriptor) in forward order."""
    spec = O.SPECS[model_type]
    lp, cd = spec["linear_proj"], spec["ctx_dim"]
    P = "model."
    yield P + "time_embedding.linear_1", ("lin", 1280, 320, True)
    yield P + "time_embedding.linear_2", ("lin", 1280, 1280, True)
    if model_type == "sdxl":
        yield P + "add_embedding.linear_1", ("lin", 1280, 2816, True)
        yield P + "add_embedding.linear_2", ("lin", 1280, 1280, True)
    yield P + "conv_in", ("conv", 320, 4, 3)
    skip_ch = [320]
    for i, (cin, cout, nl, has_down) in enumerate(spec["down"]):
        for j in range(2):
            rin = cin if j == 0 else cout
            # sd.py:511-514: block 0 has no shortcut (in == out); sdxl always passes conv_shortcut=True
            # for CrossAttnDownBlock2D resnet 0 and False elsewhere (sdxl.py:367-397)
            if model_type == "sd":
                sc = (j == 0 and cin != cout)
            else:
                sc = (j == 0 and nl is not None)
            yield from _resnet(f"{P}down_blocks.{i}.resnets.{j}", rin, cout, sc)
            if nl is not None:
                yield from _t2d(f"{P}down_blocks.{i}.attentions.{j}", cout, nl, cd, lp)
            skip_ch.append(cout)
        if has_down:
            yield f"{P}down_blocks.{i}.downsamplers.0.conv", ("conv", cout, cout, 3)
            skip_ch.append(cout)
    yield from _resnet(P + "mid_block.resnets.0", 1280, 1280, False)
    yield from _t2d(P + "mid_block.attentions.0", 1280, spec["mid_layers"], cd, lp)
    yield from _resnet(P + "mid_block.resnets.1", 1280, 1280, False)
    h = 1280
    for i, (cin, cout, prev, nl, has_up) in enumerate(spec["up"]):
        for j in range(3):
            rin = h + skip_ch.pop()
            yield from _resnet(f"{P}up_blocks.{i}.resnets.{j}", rin, cout, True)
            h = cout
            if nl is not None:
                yield from _t2d(f"{P}up_blocks.{i}.attentions.{j}", cout, nl, cd, lp)
        if has_up:
            yield f"{P}up_blocks.{i}.upsamplers.0.conv", ("conv", cout, cout, 3)
    yield P + "conv_norm_out", ("gn", 320)
    yield P + "conv_out", ("conv", 4, 320, 3)


def _rng(seed: int, name: str) -> np.random.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return np.random.Generator(np.random.PCG64(int.from_bytes(h[:8], "little")))


def make_weights(model_type: str, seed: int = 0) -> Dict[str, Tensor]:
    """Random-init weights with nn.Conv2d/nn.Linear-like scale (uniform
    +-1/sqrt(fan_in)); norm affine parameters perturbed off (1, 0)."""
    sd: Dict[str, Tensor] = {}
    for name, d in iter_modules(model_type):
        g = _rng(seed, name)
        if d[0] in ("gn", "ln"):
            c = d[1]
            sd[name + ".weight"] = torch.from_numpy((1.0 + 0.1 * g.standard_normal(c)).astype(np.float32))
            sd[name + ".bias"] = torch.from_numpy((0.1 * g.standard_normal(c)).astype(np.float32))
        elif d[0] == "conv":
            _, co, ci, k = d
            bound = 1.0 / math.sqrt(ci * k * k)
            sd[name + ".w"] = torch.from_numpy(g.uniform(-bound, bound, (co, ci, k, k)).astype(np.float32))
            sd[name + ".b"] = torch.from_numpy(g.uniform(-bound, bound, (co,)).astype(np.float32))
        else:
            _, n, k, bias = d
            bound = 1.0 / math.sqrt(k)
            sd[name + ".w"] = torch.from_numpy(g.uniform(-bound, bound, (n, k)).astype(np.float32))
            if bias:
                sd[name + ".b"] = torch.from_numpy(g.uniform(-bound, bound, (n,)).astype(np.float32))
    return sd


def init_weight_quant(sd: Dict[str, Tensor], wbits: int, adaround: bool = False, seed: int = 0) -> None:
    """Add ``.wqtizer.delta/.zero_point`` (and optionally ``.alpha``) in place.
    Vectorised form of channel_minmax_scale; same float64 -> fp32 path as the
    reference's python-float arithmetic (quant/quant_layer.py:22-38)."""
    level = 2 ** wbits
    for key in [k for k in sd if k.endswith(".w")]:
        name = key[:-2]
        w = sd[key]
        flat = w.reshape(w.shape[0], -1).double()
        lo = torch.clamp(flat.min(dim=1)[0], max=0.0)
        hi = torch.clamp(flat.max(dim=1)[0], min=0.0)
        delta = ((hi - lo) / (level - 1)).float()
        delta = torch.where(delta < 1e-8, torch.full_like(delta, 1e-8), delta)
        zp = torch.round(-lo.float() / delta)
        shape = (-1,) + (1,) * (w.dim() - 1)
        sd[name + ".wqtizer.delta"] = delta.view(shape)
        sd[name + ".wqtizer.zero_point"] = zp.view(shape)
        if adaround:
            g = _rng(seed, name + ".alpha")
            sd[name + ".wqtizer.alpha"] = torch.from_numpy(
                g.standard_normal(tuple(w.shape)).astype(np.float32))


# --------------------------------------------------------------------------- #
# activation scales
# --------------------------------------------------------------------------- #
class Recorder(dict):
    """Passed as ``act`` to the oracle with ``cfg.use_aq`` on: records, per
    quantizer key, the min/max statistics the reference's record_min_max_ema
    keeps (quant/quant_layer.py:301-313), and quantizes nothing."""

    def __init__(self):
        super().__init__()
        self.stats: Dict[str, dict] = {}

    def __contains__(self, key):  # every quantizer "exists"
        return isinstance(key, str) and key.endswith(".delta")

    def observe(self, key: str, x: Tensor) -> None:
        st = {"dim": x.dim(), "min": float(x.min()), "max": float(x.max())}
        if x.dim() == 3:
            st["in"] = (x.amin(dim=(0, 1)), x.amax(dim=(0, 1)))
            st["out"] = (x.amin(dim=(0, 2)), x.amax(dim=(0, 2)))
        elif x.dim() == 4:
            st["in"] = (x.amin(dim=(0, 1, 2)), x.amax(dim=(0, 1, 2)))
            st["out"] = (x.amin(dim=(0, 1, 3)), x.amax(dim=(0, 1, 3)))
        self.stats[key] = st


def _group_scales(lo: Tensor, hi: Tensor, level: int, g: int) -> Tuple[Tensor, Tensor]:
    """Cluster channels by range into ``g`` non-contiguous groups; per-group
    (delta, zp) from the cluster's min/max ('minmax' mode, quant_layer.py:380-423)."""
    n = lo.numel()
    order = torch.argsort(hi - lo, stable=True)
    labels = torch.empty(n, dtype=torch.long)
    labels[order] = torch.arange(n) * g // n
    delta = torch.empty(n)
    zp = torch.empty(n)
    for i in range(g):
        m = labels == i
        if not m.any():
            continue
        cmin = float(min(lo[m].min(), hi[m].min()))
        cmax = float(max(lo[m].max(), hi[m].max()))
        d = torch.tensor((cmax - cmin) / (level - 1))
        if d < 1e-8:
            d = torch.tensor(1e-8)
        delta[m] = d
        zp[m] = torch.round(torch.tensor(-cmin) / d)
    return delta, zp


def calibrate_act(model_type: str, sd, cfg: O.QConfig, inputs: tuple, group_num: int,
                  force: Optional[str] = None) -> Dict[str, Tensor]:
    """One ``act_k`` dict from a recording FP-activation forward of the oracle.

    ``force``: None = reference's spread heuristic (quant_layer.py:345-352),
    'in' = always (1,1,X) (env IN_CHANNEL_WISE), 'out' = always (1,X,1)."""
    rec = Recorder()
    hooked = _HookedOracle(rec)
    with hooked:
        O.unet_forward(model_type, sd, rec, cfg, *inputs)
    act: Dict[str, Tensor] = {}
    level = 2 ** cfg.abits
    for key, st in rec.stats.items():
        if key.endswith("aqtizer_w"):
            lv = 2 ** cfg.softmax_bits
            if cfg.t2i_log_quant:
                continue  # T2ILogQuantizer saves nothing; real-time only
            act[key + ".delta"] = torch.tensor(max(st["max"], 0.0) / (lv - 1))
            act[key + ".zero_point"] = torch.tensor(0.0)
            continue
        if st["dim"] <= 2 or group_num <= 1:
            lo, hi = min(st["min"], 0.0), max(st["max"], 0.0)
            d = torch.tensor(float(hi - lo) / (level - 1))
            if d < 1e-8:
                d = torch.tensor(1e-8)
            act[key + ".delta"] = d
            act[key + ".zero_point"] = torch.round(torch.tensor(-lo) / d)
            continue
        (ilo, ihi), (olo, ohi) = st["in"], st["out"]
        spread_in = float(ihi.max() - ihi.min() + ilo.max() - ilo.min())
        spread_out = float(ohi.max() - ohi.min() + olo.max() - olo.min())
        use_in = spread_in > spread_out if force is None else force == "in"
        if use_in:
            d, z = _group_scales(ilo, ihi, level, group_num)
            act[key + ".delta"], act[key + ".zero_point"] = d.view(1, 1, -1), z.view(1, 1, -1)
        else:
            d, z = _group_scales(olo, ohi, level, group_num)
            act[key + ".delta"], act[key + ".zero_point"] = d.view(1, -1, 1), z.view(1, -1, 1)
    return act


class _HookedOracle:
    """Temporarily reroutes the oracle's quantizer calls into a Recorder."""

    def __init__(self, rec: Recorder):
        self.rec = rec

    def __enter__(self):
        self._uaq = O.uaq_fake_quant
        self._aq = O._aq
        self._ql = O.quant_layer
        rec = self.rec
        orig_ql = self._ql

        def aq(act, key, x, level):
            rec.observe(key, x)
            return x

        def ql(x, sd, act, name, cfg, *, stride=1, padding=0, fp_layer=False):
            if not fp_layer:
                w = sd[name + ".w"]
                xo = x
                if w.dim() == 4:
                    xo = torch.nn.functional.unfold(x, kernel_size=(w.shape[2], w.shape[3]),
                                                    padding=padding, stride=stride)
                rec.observe(name + ".aqtizer", xo)
            return orig_ql(x, sd, None, name, cfg, stride=stride, padding=padding, fp_layer=fp_layer)

        def mq(m, act, name, cfg):  # softmax-map quantizer: record only
            rec.observe(name + ".aqtizer_w", m)
            return m

        self._mq = O._map_quant
        O._map_quant = mq
        O._aq = aq
        O.quant_layer = ql
        return self

    def __exit__(self, *exc):
        O._aq = self._aq
        O.quant_layer = self._ql
        O._map_quant = self._mq
        return False


def random_act(model_type: str, sd, cfg: O.QConfig, shapes: Dict[str, Tuple[str, int]],
               group_num: int, seed: int) -> Dict[str, Tensor]:
    """SURVEY.md 8d synthetic scales: per quantizer, labels ~ randint(0,g),
    lo = -(U*3+1), hi = U*3+1.  ``shapes[key] = (orientation, X)`` with
    orientation in {'scalar','in','out'}."""
    level = 2 ** cfg.abits
    act: Dict[str, Tensor] = {}
    for key, (orient, n) in shapes.items():
        g = _rng(seed, key)
        if key.endswith("aqtizer_w"):
            act[key + ".delta"] = torch.tensor(1.0 / (2 ** cfg.softmax_bits - 1))
            act[key + ".zero_point"] = torch.tensor(0.0)
            continue
        ng = 1 if orient == "scalar" else group_num
        lo = -(g.uniform(0, 1, ng) * 3 + 1)
        hi = g.uniform(0, 1, ng) * 3 + 1
        d = ((hi - lo) / (level - 1)).astype(np.float32)
        z = np.round((-lo).astype(np.float32) / d)
        if orient == "scalar":
            act[key + ".delta"] = torch.tensor(float(d[0]))
            act[key + ".zero_point"] = torch.tensor(float(z[0]))
            continue
        labels = g.integers(0, ng, n)
        dt, zt = torch.from_numpy(d[labels]), torch.from_numpy(z[labels])
        view = (1, 1, -1) if orient == "in" else (1, -1, 1)
        act[key + ".delta"], act[key + ".zero_point"] = dt.view(view), zt.view(view)
    return act


def example_inputs(model_type: str, batch: int, seed: int = 0, t: int = 500):
    """Synthetic UNet inputs of SURVEY.md 8d (torch CPU generator, seeded)."""
    g = torch.Generator().manual_seed(seed)
    spec = O.SPECS[model_type]
    s = spec["sample"]
    sample = torch.randn(batch, 4, s, s, generator=g)
    ctx = torch.randn(batch, 77, spec["ctx_dim"], generator=g)
    ts = torch.tensor([t])
    if model_type == "sdxl":
        added = {"text_embeds": torch.randn(batch, 1280, generator=g),
                 "time_ids": torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]]).repeat(batch, 1)}
        return sample, ts, ctx, added
    return sample, ts, ctx
