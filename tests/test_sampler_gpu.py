"""Device sampler (dgq_sampler_step + dgq_b200/sampler.py) against the reference schedulers' golden trajectories
(tests/golden/sampler.pt), against the CPU oracle loops, and end to end around the quantized SD UNet."""
import os

import pytest
import torch

from oracle import dgq_oracle as O, sampler_oracle as SO, synth as S
from tests import unet_cases as U

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler.pt"))
SD_CFG = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1)
XL_CFG = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", timestep_spacing="trailing")


def dummy_sample():
    n = 4 * 3 * 8 * 8
    return (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2).contiguous()


@pytest.mark.parametrize("key,n,cfg", [("pndm_sd_50", 50, SD_CFG), ("pndm_sd_10", 10, SD_CFG),
                                       ("pndm_sd_vpred_10", 10, dict(prediction_type="v_prediction", **SD_CFG))])
def test_plms_matches_reference_trajectory(key, n, cfg):
    from dgq_b200.sampler import PLMSSampler
    g = GOLD[key]
    s = PLMSSampler(**cfg)
    s.set_timesteps(n)
    assert s.timesteps.tolist() == g["timesteps"].tolist()
    x = dummy_sample().to(DEV)
    for k, t in enumerate(s.timesteps):
        out = (x * float(t) / (float(t) + 1)).contiguous()
        x = s.step(out, int(t), x)
        assert torch.allclose(x.cpu(), g["traj"][k], rtol=2e-5, atol=2e-6), k


@pytest.mark.parametrize("key,cfg", [("euler_test_10", dict(num_train_timesteps=1100, beta_start=0.0001, beta_end=0.02,
                                                             beta_schedule="linear", timestep_spacing="linspace")),
                                     ("euler_vpred_10", dict(num_train_timesteps=1100, beta_start=0.0001, beta_end=0.02,
                                                             beta_schedule="linear", timestep_spacing="linspace",
                                                             prediction_type="v_prediction")),
                                     ("euler_turbo_1", XL_CFG), ("euler_turbo_4", XL_CFG)])
def test_euler_ancestral_matches_reference_trajectory(key, cfg):
    from dgq_b200.sampler import EulerAncestralSampler
    g = GOLD[key]
    s = EulerAncestralSampler(**cfg)
    s.set_timesteps(len(g["timesteps"]))
    assert s.timesteps.tolist() == g["timesteps"].tolist()
    assert abs(s.init_noise_sigma - g["init_noise_sigma"].item()) < 1e-6 * g["init_noise_sigma"].item()
    x = (dummy_sample() * s.init_noise_sigma).to(DEV)
    model_in = (x * s.input_scale(0)).contiguous()
    for k, t in enumerate(s.timesteps):
        out = (model_in * float(t) / (float(t) + 1)).contiguous()
        x = s.step(out, t, x, g["noises"][k].to(DEV).contiguous(), model_in=model_in)
        assert torch.allclose(x.cpu(), g["traj"][k], rtol=3e-5, atol=3e-5), k


def test_cfg_combine_and_model_input():
    from dgq_b200.sampler import sampler_step
    g = torch.Generator().manual_seed(0)
    u, c, x = (torch.randn(3, 4, 8, 8, generator=g) for _ in range(3))
    out = torch.empty(3, 4, 8, 8, device=DEV)
    eps = torch.empty_like(out)
    mi = torch.empty(6, 4, 8, 8, device=DEV)
    sampler_step(torch.cat([u, c]).to(DEV), x.to(DEV), out, cx=0.9, c_eps=-0.3, guidance=7.5, eps_store=eps,
                 model_in=mi, in_scale=0.5)
    want_eps = u + 7.5 * (c - u)
    assert torch.equal(eps.cpu(), want_eps)                      # same operation order as the pipeline: bit-exact
    want = 0.9 * x - 0.3 * want_eps
    assert torch.allclose(out.cpu(), want, rtol=1e-6, atol=1e-6)
    assert torch.equal(mi[:3], mi[3:]) and torch.allclose(mi[:3].cpu(), want * 0.5, rtol=1e-6, atol=1e-6)


def test_denoise_loops_match_oracle_with_a_stand_in_unet():
    """the loops' wiring (CFG batch, timesteps handed to the UNet, model-input scaling) with a cheap UNet stand-in"""
    from dgq_b200 import sampler as DS
    g = torch.Generator().manual_seed(1)
    lat = torch.randn(2, 4, 16, 16, generator=g)
    ctx = torch.randn(4, 77, 8, generator=g)

    def fake(x, t, c, *a, **k):
        bias = c.float().mean(dim=(1, 2)).view(-1, 1, 1, 1).to(x.device)
        return [torch.tanh(x * 0.7 + bias) * (1.0 + float(t.reshape(-1)[0]) / 1000.0)]
    want = SO.denoise_sd(lambda x, t, c: fake(x, t, c)[0], lat, ctx, 10, guidance=7.5)
    got = DS.denoise_sd(fake, lat.to(DEV), ctx.to(DEV), 10, guidance=7.5)
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-4)
    noises = [torch.randn(2, 4, 16, 16, generator=g) for _ in range(4)]
    want = SO.denoise_sdxl(lambda x, t, c, a: fake(x, t, c)[0], lat, ctx[:2], {}, 4, noises)
    got = DS.denoise_sdxl(fake, lat.to(DEV), ctx[:2].to(DEV), {}, 4, [n.to(DEV) for n in noises])
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-4)


def test_sd_cfg_plms_loop_end_to_end(tmp_path):
    """BASELINE config 3 in small: SD W4A8 g8 + t2i log (real-time, start-peak), time-aware, PLMS with CFG,
    through get_qmodel + the device sampler vs the CPU oracle UNet + oracle sampler (2 steps = 3 UNet calls).
    Two 500-timestep PLMS steps with guidance 7.5 are the harshest setting there is: CFG multiplies the
    difference of two UNet outputs by 7.5 and each step moves the latent by O(1).  scripts/sd_loop_parity.py runs
    longer loops (10 steps: see DESIGN.md "parity").  Bar: 0.998 (per-call bar of tests/test_unet_gpu.py)."""
    from dgq_b200 import sampler as DS
    model_type, case = "sd", "w4a8_g8_log"
    sd, cfg, acts = U.build_case(S, O, model_type, case, torch)
    qnn = U.build_qmodel(model_type, case, sd, acts, tmp_path)
    g = torch.Generator().manual_seed(7)
    lat = torch.randn(1, 4, 64, 64, generator=g)
    ctx = torch.randn(2, 77, 768, generator=g)
    n_steps = len(acts)

    def oracle_unet(x, t, c):
        idx = int((1000 - float(t)) // (1000 // n_steps))
        O.update_group_convs(cfg, acts[idx], sd)
        return O.unet_forward(model_type, sd, acts[idx], cfg, x, t, c)
    with torch.no_grad():
        want = SO.denoise_sd(oracle_unet, lat, ctx, n_steps, guidance=7.5)
        got = DS.denoise_sd(qnn, lat.to(DEV), ctx.to(DEV), n_steps, guidance=7.5)
    cos = U.cosine(got, want)
    print(f"sd CFG PLMS {n_steps}-step loop: final-latent cosine {cos:.6f}")
    assert cos >= 0.998, cos
