"""CPU oracle for the sampler loop around the quantized UNet -- TEST INFRASTRUCTURE ONLY.

Restates, in plain torch-CPU fp32, the two schedulers the reference's pipelines drive the hot path
with, and the denoise loops that call the UNet:

  PNDM / PLMS          diffusers/src/diffusers/schedulers/scheduling_pndm.py:168-226 (set_timesteps),
                       :262-319 (step_prk), :321-387 (step_plms), :407-449 (_get_prev_sample)
  Euler ancestral      .../scheduling_euler_ancestral_discrete.py:262-305 (set_timesteps),
                       :239-260 (scale_model_input), :323-414 (step)
  SD denoise loop      .../pipelines/stable_diffusion/pipeline_stable_diffusion.py:1017-1047
  SDXL denoise loop    .../pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl.py:1234-1267

Pinned by the reference's own known-answer tests (tests/test_sampler_cpu.py):
  diffusers/tests/schedulers/test_scheduler_pndm.py:210-224  (sum 198.1318 / mean 0.2580; v-pred 67.3986 / 0.0878)
  diffusers/tests/schedulers/test_scheduler_euler_ancestral.py:44-99 (152.3192 / 0.1983; v-pred 108.4439 / 0.1412)
Nothing under dgq_b200/ imports this module.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np
import torch


def make_betas(schedule: str, beta_start: float, beta_end: float, n: int) -> torch.Tensor:
    if schedule == "linear":
        return torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    if schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    raise NotImplementedError(schedule)


class PNDMOracle:
    """F-PNDM (4 Runge-Kutta warm-up steps x 3, then linear multistep) or, with skip_prk_steps, PLMS."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 skip_prk_steps=False, set_alpha_to_one=False, prediction_type="epsilon", steps_offset=0):
        self.T = num_train_timesteps
        self.acp = torch.cumprod(1.0 - make_betas(beta_schedule, beta_start, beta_end, num_train_timesteps), 0)
        self.final_acp = torch.tensor(1.0) if set_alpha_to_one else self.acp[0]
        self.skip_prk, self.pred, self.offset = skip_prk_steps, prediction_type, steps_offset
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n: int) -> None:
        self.n = n
        ratio = self.T // n
        base = (np.arange(0, n) * ratio).round() + self.offset          # "leading" spacing
        if self.skip_prk:
            self.prk_t = np.array([])
            self.plms_t = np.concatenate([base[:-1], base[-2:-1], base[-1:]])[::-1].copy()
        else:
            prk = np.array(base[-4:]).repeat(2) + np.tile(np.array([0, self.T // n // 2]), 4)
            self.prk_t = (prk[:-1].repeat(2)[1:-1])[::-1].copy()
            self.plms_t = base[:-3][::-1].copy()
        self.timesteps = np.concatenate([self.prk_t, self.plms_t]).astype(np.int64)
        self.ets: List[torch.Tensor] = []
        self.counter = 0
        self.acc = 0
        self.held = None

    def _prev(self, x, t, t_prev, eps):
        a_t = self.acp[t]
        a_p = self.acp[t_prev] if t_prev >= 0 else self.final_acp
        b_t, b_p = 1 - a_t, 1 - a_p
        if self.pred == "v_prediction":
            eps = (a_t ** 0.5) * eps + (b_t ** 0.5) * x
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return (a_p / a_t) ** 0.5 * x - (a_p - a_t) * eps / denom

    def step(self, eps, t, x):
        if self.counter < len(self.prk_t) and not self.skip_prk:
            return self.step_prk(eps, t, x)
        return self.step_plms(eps, t, x)

    def step_prk(self, eps, t, x):
        t = int(t)
        half = 0 if self.counter % 2 else self.T // self.n // 2
        t_prev = t - half
        t0 = int(self.prk_t[self.counter // 4 * 4])
        k = self.counter % 4
        if k == 0:
            self.acc = self.acc + eps / 6
            self.ets.append(eps)
            self.held = x
        elif k in (1, 2):
            self.acc = self.acc + eps / 3
        else:
            eps = self.acc + eps / 6
            self.acc = 0
        out = self._prev(self.held if self.held is not None else x, t0, t_prev, eps)
        self.counter += 1
        return out

    def step_plms(self, eps, t, x):
        t = int(t)
        if not self.skip_prk and len(self.ets) < 3:
            raise ValueError("PLMS needs the Runge-Kutta warm-up unless skip_prk_steps is set")
        step = self.T // self.n
        t_prev = t - step
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(eps)
        else:
            t_prev, t = t, t + step
        e = self.ets
        if len(e) == 1 and self.counter == 0:
            self.held = x
        elif len(e) == 1 and self.counter == 1:
            eps = (eps + e[-1]) / 2
            x, self.held = self.held, None
        elif len(e) == 2:
            eps = (3 * e[-1] - e[-2]) / 2
        elif len(e) == 3:
            eps = (23 * e[-1] - 16 * e[-2] + 5 * e[-3]) / 12
        else:
            eps = (55 * e[-1] - 59 * e[-2] + 37 * e[-3] - 9 * e[-4]) / 24
        out = self._prev(x, t, t_prev, eps)
        self.counter += 1
        return out

    def scale_model_input(self, x, t=None):
        return x


class EulerAncestralOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", timestep_spacing="linspace", steps_offset=0):
        self.T = num_train_timesteps
        self.acp = torch.cumprod(1.0 - make_betas(beta_schedule, beta_start, beta_end, num_train_timesteps), 0)
        self.pred, self.spacing, self.offset = prediction_type, timestep_spacing, steps_offset
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, n: int) -> None:
        if self.spacing == "linspace":
            ts = np.linspace(0, self.T - 1, n, dtype=np.float32)[::-1].copy()
        elif self.spacing == "leading":
            ts = (np.arange(0, n) * (self.T // n)).round()[::-1].copy().astype(np.float32) + self.offset
        elif self.spacing == "trailing":
            ts = (np.arange(self.T, 0, -self.T / n)).round().copy().astype(np.float32) - 1
        else:
            raise ValueError(self.spacing)
        sig = (((1 - self.acp) / self.acp) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.i = 0

    @property
    def init_noise_sigma(self):
        m = self.sigmas.max()
        return m if self.spacing in ("linspace", "trailing") else (m ** 2 + 1) ** 0.5

    def scale_model_input(self, x, t=None):
        s = self.sigmas[self.i]
        return x / ((s ** 2 + 1) ** 0.5)

    def step(self, out, t, x, noise):
        """`noise` = the N(0,1) tensor the reference draws with randn_tensor(generator)."""
        s, s_to = self.sigmas[self.i], self.sigmas[self.i + 1]
        x = x.to(torch.float32)
        if self.pred == "epsilon":
            x0 = x - s * out
        elif self.pred == "v_prediction":
            x0 = out * (-s / (s ** 2 + 1) ** 0.5) + (x / (s ** 2 + 1))
        else:
            raise ValueError(self.pred)
        up = (s_to ** 2 * (s ** 2 - s_to ** 2) / s ** 2) ** 0.5
        down = (s_to ** 2 - up ** 2) ** 0.5
        nxt = x + (x - x0) / s * (down - s) + noise * up
        self.i += 1
        return nxt.to(out.dtype)


# --------------------------------------------------------------------------------------------
def denoise_sd(unet: Callable, latents: torch.Tensor, ctx_uncond_cond: torch.Tensor, n_steps: int,
               guidance: float = 7.5, sched: Optional[PNDMOracle] = None) -> torch.Tensor:
    """StableDiffusionPipeline.__call__ denoise loop (pipeline_stable_diffusion.py:1017-1047) with
    SD-v1.4's stock PNDM config; `unet(x, t, ctx) -> noise`, ctx = cat([uncond, cond])."""
    s = sched or PNDMOracle(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", skip_prk_steps=True,
                            steps_offset=1)
    s.set_timesteps(n_steps)
    x = latents * s.init_noise_sigma
    cfg = guidance > 1.0
    for t in s.timesteps:
        xin = torch.cat([x] * 2) if cfg else x
        out = unet(s.scale_model_input(xin, t), torch.tensor([float(t)]), ctx_uncond_cond)
        if cfg:
            u, c = out.chunk(2)
            out = u + guidance * (c - u)
        x = s.step(out, t, x)
    return x


def denoise_sdxl(unet: Callable, latents: torch.Tensor, ctx: torch.Tensor, added: dict, n_steps: int,
                 noises: List[torch.Tensor], sched: Optional[EulerAncestralOracle] = None) -> torch.Tensor:
    """StableDiffusionXLPipeline denoise loop for SDXL-turbo (guidance 0 => no CFG,
    pipeline_stable_diffusion_xl.py:1234-1267), EulerAncestral with trailing spacing."""
    s = sched or EulerAncestralOracle(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                      timestep_spacing="trailing")
    s.set_timesteps(n_steps)
    x = latents * s.init_noise_sigma
    for k, t in enumerate(s.timesteps):
        out = unet(s.scale_model_input(x, t), torch.tensor([float(t)]), ctx, added)
        x = s.step(out, t, x, noises[k])
    return x
