"""Device-resident sampler loops around the quantized UNet.

The reference drives the UNet from diffusers' pipelines: per step a python scheduler object does a
dozen small eager tensor ops (CFG chunk/combine, alpha lookups on CPU tensors, the multistep
combination, a fresh `torch.cat([latents] * 2)`), see pipeline_stable_diffusion.py:1017-1047 and
pipeline_stable_diffusion_xl.py:1234-1267.  Every one of those updates is a linear map of
(sample, model outputs, noise), so here the host only turns the scheduler state into a handful of
fp32 coefficients (computed in float64 from the same fp32 alpha table) and ONE kernel
(`dgq_sampler_step`) applies CFG, the update and the next model input in a single pass.

  PLMSSampler            = PNDMScheduler(skip_prk_steps=True)   schedulers/scheduling_pndm.py:168-226,321-449
  EulerAncestralSampler  = EulerAncestralDiscreteScheduler      schedulers/scheduling_euler_ancestral_discrete.py:239-414
  denoise_sd / denoise_sdxl = the pipelines' denoise loops

F-PNDM's Runge-Kutta warm-up (skip_prk_steps=False) is not used by SD v1.4 / DGQ and is not built.
Tensors are CUDA fp32; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import ops


def _betas(schedule: str, beta_start: float, beta_end: float, n: int) -> torch.Tensor:
    if schedule == "linear":
        return torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    if schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    raise NotImplementedError(f"beta_schedule {schedule!r}")


def sampler_step(unet_out: torch.Tensor, x: torch.Tensor, out: torch.Tensor, *, cx: float, c_eps: float,
                 guidance: Optional[float] = None, eps_store: Optional[torch.Tensor] = None,
                 hist: Sequence[torch.Tensor] = (), c_hist: Sequence[float] = (),
                 noise: Optional[torch.Tensor] = None, c_noise: float = 0.0,
                 model_in: Optional[torch.Tensor] = None, in_scale: float = 1.0) -> torch.Tensor:
    """One launch of dgq_sampler_step (see include/dgq_b200.h)."""
    for t in (unet_out, x, out):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("sampler_step: CUDA fp32 contiguous tensors only (no CPU fallback)")
    n = x.numel()
    use_cfg = guidance is not None
    assert unet_out.numel() == (2 * n if use_cfg else n) and out.numel() == n and len(hist) == len(c_hist) <= 4
    dup = model_in is not None and model_in.numel() == 2 * n
    hp = (C.c_void_p * 4)(*[h.data_ptr() for h in hist], *([None] * (4 - len(hist))))
    hc = (C.c_float * 4)(*[float(c) for c in c_hist], *([0.0] * (4 - len(hist))))
    a = L.SamplerStepT(unet_out.data_ptr(), n, float(guidance or 0.0), int(use_cfg),
                       None if eps_store is None else eps_store.data_ptr(), x.data_ptr(), float(cx), float(c_eps),
                       hp, hc, None if noise is None else noise.data_ptr(), float(c_noise), out.data_ptr(),
                       None if model_in is None else model_in.data_ptr(), float(in_scale), int(dup))
    L.check(L.lib().dgq_sampler_step(C.byref(a), torch.cuda.current_stream().cuda_stream), "dgq_sampler_step")
    ops._count()
    return out


class PLMSSampler:
    """PNDMScheduler with skip_prk_steps=True (SD v1.4's scheduler): pseudo linear multistep."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", set_alpha_to_one: bool = False,
                 prediction_type: str = "epsilon", steps_offset: int = 1, skip_prk_steps: bool = True):
        if not skip_prk_steps:
            raise NotImplementedError("F-PNDM Runge-Kutta warm-up: SD v1.4 / DGQ run PNDM with skip_prk_steps=True")
        if prediction_type not in ("epsilon", "v_prediction"):
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon` or `v_prediction`")
        self.T = num_train_timesteps
        acp = torch.cumprod(1.0 - _betas(beta_schedule, beta_start, beta_end, num_train_timesteps), 0)
        self.acp = acp.double().numpy()
        self.final_acp = 1.0 if set_alpha_to_one else float(self.acp[0])
        self.pred, self.offset = prediction_type, steps_offset
        self.init_noise_sigma = 1.0
        self.timesteps = None

    def set_timesteps(self, n: int) -> None:
        self.n = n
        base = (np.arange(0, n) * (self.T // n)).round() + self.offset
        self.timesteps = np.concatenate([base[:-1], base[-2:-1], base[-1:]])[::-1].astype(np.int64).copy()
        self.counter = 0
        self._ets: List[torch.Tensor] = []      # newest last
        self._free: List[torch.Tensor] = []
        self._held = None
        self._bufs = None

    def _coef(self, t: int, t_prev: int):
        """prev = cx * X + cm * m   (m = the multistep combination of model outputs)"""
        a_t = self.acp[t]
        a_p = self.acp[t_prev] if t_prev >= 0 else self.final_acp
        b_t, b_p = 1 - a_t, 1 - a_p
        A = (a_p / a_t) ** 0.5
        B = (a_p - a_t) / (a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5)
        if self.pred == "v_prediction":     # m' = sqrt(a_t) m + sqrt(b_t) X
            return A - B * b_t ** 0.5, -B * a_t ** 0.5
        return A, -B

    def _alloc(self, like: torch.Tensor):
        if self._bufs is None:
            self._bufs = [torch.empty_like(like) for _ in range(2)]
            self._free = [torch.empty_like(like) for _ in range(4)]
            self._held = torch.empty_like(like)
            self._flip = 0

    def step(self, unet_out: torch.Tensor, t: int, x: torch.Tensor, *, guidance: Optional[float] = None,
             model_in: Optional[torch.Tensor] = None) -> torch.Tensor:
        """unet_out: UNet output ([2N,...] with guidance, uncond first); x: current latents [N,...].
        Returns the previous-timestep latents; optionally writes the next model input (CFG-duplicated)."""
        if self.timesteps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = int(t)
        self._alloc(x)
        step = self.T // self.n
        t_prev = t - step
        store = None
        if self.counter != 1:
            store = self._free.pop() if self._free else self._ets.pop(0)
        else:
            t_prev, t = t, t + step
        e = self._ets
        k = len(e) + (1 if store is not None else 0)
        X = x
        if self.counter == 0:
            self._held.copy_(x)
            w_new, hist, w_hist = 1.0, [], []
        elif self.counter == 1:
            X = self._held
            w_new, hist, w_hist = 0.5, [e[-1]], [0.5]
        elif k == 2:
            w_new, hist, w_hist = 1.5, [e[-1]], [-0.5]
        elif k == 3:
            w_new, hist, w_hist = 23 / 12, [e[-1], e[-2]], [-16 / 12, 5 / 12]
        else:
            w_new, hist, w_hist = 55 / 24, [e[-1], e[-2], e[-3]], [-59 / 24, 37 / 24, -9 / 24]
        cx, cm = self._coef(t, t_prev)
        out = self._bufs[self._flip]
        self._flip ^= 1
        sampler_step(unet_out, X, out, cx=cx, c_eps=cm * w_new, guidance=guidance, eps_store=store, hist=hist,
                     c_hist=[cm * w for w in w_hist], model_in=model_in, in_scale=1.0)
        if store is not None:
            e.append(store)
            if len(e) > 4:
                self._free.append(e.pop(0))
        self.counter += 1
        return out

    def scale_model_input(self, x, t=None):
        return x


class EulerAncestralSampler:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", prediction_type: str = "epsilon",
                 timestep_spacing: str = "trailing", steps_offset: int = 0):
        if prediction_type not in ("epsilon", "v_prediction"):
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon`, or `v_prediction`")
        self.T = num_train_timesteps
        self.acp = torch.cumprod(1.0 - _betas(beta_schedule, beta_start, beta_end, num_train_timesteps), 0)
        self.pred, self.spacing, self.offset = prediction_type, timestep_spacing, steps_offset
        self.timesteps = None
        self._bufs = None

    def set_timesteps(self, n: int) -> None:
        if self.spacing == "linspace":
            ts = np.linspace(0, self.T - 1, n, dtype=np.float32)[::-1].copy()
        elif self.spacing == "leading":
            ts = (np.arange(0, n) * (self.T // n)).round()[::-1].copy().astype(np.float32) + self.offset
        elif self.spacing == "trailing":
            ts = (np.arange(self.T, 0, -self.T / n)).round().copy().astype(np.float32) - 1
        else:
            raise ValueError(f"{self.spacing} is not supported. Please make sure to choose one of 'linspace', 'leading' or 'trailing'.")
        sig = (((1 - self.acp) / self.acp) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32).astype(np.float64)
        self.timesteps = ts
        self.i = 0

    @property
    def init_noise_sigma(self) -> float:
        m = float(self.sigmas.max())
        return m if self.spacing in ("linspace", "trailing") else (m ** 2 + 1) ** 0.5

    def input_scale(self, i: Optional[int] = None) -> float:
        s = self.sigmas[self.i if i is None else i]
        return float(1.0 / (s ** 2 + 1) ** 0.5)

    def step(self, unet_out: torch.Tensor, t, x: torch.Tensor, noise: Optional[torch.Tensor], *,
             guidance: Optional[float] = None, model_in: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.timesteps is None:
            raise ValueError("run set_timesteps first")
        if self._bufs is None or self._bufs[0].shape != x.shape:
            self._bufs, self._flip = [torch.empty_like(x) for _ in range(2)], 0
        s, s_to = self.sigmas[self.i], self.sigmas[self.i + 1]
        up = (s_to ** 2 * (s ** 2 - s_to ** 2) / s ** 2) ** 0.5
        down = (s_to ** 2 - up ** 2) ** 0.5
        dt = down - s
        if self.pred == "epsilon":              # derivative = eps
            cx, ce = 1.0, dt
        else:                                   # derivative = (x - x0) / s with x0 = -s/sqrt(s^2+1) out + x/(s^2+1)
            cx, ce = 1.0 + dt * (s / (s ** 2 + 1)), dt / (s ** 2 + 1) ** 0.5
        out = self._bufs[self._flip]
        self._flip ^= 1
        nxt_scale = float(1.0 / (s_to ** 2 + 1) ** 0.5)
        sampler_step(unet_out, x, out, cx=cx, c_eps=ce, guidance=guidance, noise=noise if up != 0.0 else None,
                     c_noise=float(up), model_in=model_in, in_scale=nxt_scale)
        self.i += 1
        return out


# ------------------------------------------------------------------------------------------
def denoise_sd(unet: Callable, latents: torch.Tensor, ctx_uncond_cond: torch.Tensor, n_steps: int,
               guidance: float = 7.5, sampler: Optional[PLMSSampler] = None) -> torch.Tensor:
    """StableDiffusionPipeline's denoise loop (pipeline_stable_diffusion.py:1017-1047): 50 steps => 51 UNet calls,
    CFG pair co-resident in one UNet batch.  `unet(x, t, ctx)[0]` is the QuantModel call."""
    s = sampler or PLMSSampler()
    s.set_timesteps(n_steps)
    cfg = guidance > 1.0
    x = (latents * s.init_noise_sigma).contiguous()
    model_in = torch.cat([x] * 2) if cfg else x.clone()
    for t in s.timesteps:
        out = unet(model_in, torch.tensor([float(t)]), ctx_uncond_cond)[0]
        x = s.step(out.contiguous(), int(t), x, guidance=guidance if cfg else None, model_in=model_in)
    return x


def denoise_sdxl(unet: Callable, latents: torch.Tensor, ctx: torch.Tensor, added: dict, n_steps: int,
                 noises: Sequence[torch.Tensor], sampler: Optional[EulerAncestralSampler] = None) -> torch.Tensor:
    """StableDiffusionXLPipeline's denoise loop for SDXL-turbo (guidance 0 => no CFG,
    pipeline_stable_diffusion_xl.py:1234-1267); noises[k] = the N(0,1) draw of step k."""
    s = sampler or EulerAncestralSampler()
    s.set_timesteps(n_steps)
    x = (latents * s.init_noise_sigma).contiguous()
    model_in = (x * s.input_scale(0)).contiguous()
    for k, t in enumerate(s.timesteps):
        out = unet(model_in, torch.tensor([float(t)]), ctx, added_cond_kwargs=added)[0]
        x = s.step(out.contiguous(), t, x, noises[k], model_in=model_in)
    return x
