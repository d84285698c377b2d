"""Mint sampler golden vectors by EXECUTING THE REFERENCE's schedulers (vendored diffusers 0.26.0 under
/root/reference/diffusers/src) on the dummy model of diffusers/tests/schedulers/test_schedulers.py:300-310.

    python tests/golden/make_sampler_golden.py        # -> tests/golden/sampler.pt  (~40 KB)

The Euler-ancestral noise is drawn with the reference's own randn_tensor(generator) call; the drawn tensors are
stored next to the outputs (the CPU RNG stream differs between torch builds, so the reference's checked-in
known answer 152.3192 is not reproducible with torch 2.11 -- the reference itself gives 233.2862 here)."""
import os
import sys
import types

import huggingface_hub
huggingface_hub.cached_download = lambda *a, **k: None
for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path[:0] = ["/root/reference/diffusers/src"]
import torch  # noqa: E402
from diffusers.schedulers.scheduling_euler_ancestral_discrete import EulerAncestralDiscreteScheduler  # noqa: E402
from diffusers.schedulers.scheduling_pndm import PNDMScheduler  # noqa: E402
import diffusers.schedulers.scheduling_euler_ancestral_discrete as EA  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler.pt")


def dummy_sample():
    n = 4 * 3 * 8 * 8
    return (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2).contiguous()


def model(sample, t):
    t = t.reshape(-1, 1, 1, 1).to(sample.dtype) if isinstance(t, torch.Tensor) else float(t)
    return sample * t / (t + 1)


def run_pndm(n_steps, **cfg):
    s = PNDMScheduler(**cfg)
    s.set_timesteps(n_steps)
    x = dummy_sample()
    traj = []
    for t in s.timesteps:
        x = s.step(model(x, t), t, x).prev_sample
        traj.append(x.clone())
    return {"timesteps": s.timesteps.clone(), "final": x, "traj": torch.stack(traj)}


def run_euler(n_steps, **cfg):
    s = EulerAncestralDiscreteScheduler(**cfg)
    s.set_timesteps(n_steps)
    noises = []
    orig = EA.randn_tensor

    def rec(*a, **k):
        z = orig(*a, **k)
        noises.append(z.clone())
        return z
    EA.randn_tensor = rec
    try:
        g = torch.manual_seed(0)
        x = dummy_sample() * s.init_noise_sigma
        traj = []
        for t in s.timesteps:
            x = s.step(model(s.scale_model_input(x, t), t), t, x, generator=g).prev_sample
            traj.append(x.clone())
    finally:
        EA.randn_tensor = orig
    return {"timesteps": s.timesteps.clone(), "sigmas": s.sigmas.clone(), "noises": torch.stack(noises), "final": x,
            "traj": torch.stack(traj), "init_noise_sigma": torch.as_tensor(s.init_noise_sigma)}


def main():
    sd_cfg = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", skip_prk_steps=True, steps_offset=1)
    xl_cfg = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", timestep_spacing="trailing")
    gold = {
        "pndm_sd_50": run_pndm(50, **sd_cfg),                       # SD v1.4 stock scheduler: 51 calls
        "pndm_sd_10": run_pndm(10, **sd_cfg),
        "pndm_sd_vpred_10": run_pndm(10, prediction_type="v_prediction", **sd_cfg),
        "pndm_prk_10": run_pndm(10),                                 # F-PNDM with the Runge-Kutta warm-up
        "pndm_vpred_10": run_pndm(10, prediction_type="v_prediction"),
        "euler_test_10": run_euler(10, num_train_timesteps=1100),    # the reference's test config
        "euler_vpred_10": run_euler(10, num_train_timesteps=1100, prediction_type="v_prediction"),
        "euler_turbo_1": run_euler(1, **xl_cfg),
        "euler_turbo_4": run_euler(4, **xl_cfg),
        "torch": str(torch.__version__),
    }
    torch.save(gold, OUT)
    for k, v in gold.items():
        if isinstance(v, dict):
            print(k, v["final"].abs().sum().item())


if __name__ == "__main__":
    main()
