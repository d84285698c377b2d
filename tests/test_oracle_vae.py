"""The VAE-decode oracle (oracle/vae_oracle.py) against outputs of the reference's own AutoencoderKL and
VaeImageProcessor (tests/golden/vae.pt, minted by tests/golden/make_golden.py vae).  CPU only."""
import os

import pytest
import torch

from oracle import vae_oracle as V

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vae.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


@pytest.mark.parametrize("name", ["small_b2_16", "sd_b1_8", "sdxl_b1_8"])
def test_decode_matches_reference(gold, name):
    g = gold[name]
    cfgname, b, size, wseed, iseed = g["case"]
    cfg = V.VAE_CONFIGS[cfgname]
    sd = V.make_vae_state(cfg, wseed)
    # inputs are reproducible from the seed alone
    lat = torch.randn(b, cfg["latent_channels"], size, size, generator=torch.Generator().manual_seed(iseed)) * cfg["scaling_factor"] * 4.0
    assert torch.equal(lat, g["latents"])
    with torch.no_grad():
        img = V.decode_latents(sd, cfg, lat)
    ref = g["image"]
    assert img.shape == ref.shape == (b, 3, size * 8, size * 8)
    err = float((img - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err          # fp32 vs fp32: summation order only
    u8 = V.postprocess(img)
    assert u8.shape == g["u8"].shape and u8.dtype == torch.uint8
    # rounding to 8 bits may flip where the fp32 results differ in the last bits
    assert float((u8.int() - g["u8"].int()).abs().max()) <= 1
    assert float((u8 != g["u8"]).float().mean()) < 1e-3


def test_state_enumeration_is_the_diffusers_schema():
    sd = V.make_vae_state(V.VAE_CONFIGS["sd"], 0)
    assert len(sd) == 140
    assert sd["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"].shape == (256, 512, 1, 1)
    assert sd["decoder.mid_block.attentions.0.to_q.weight"].shape == (512, 512)
    assert sd["decoder.conv_out.weight"].shape == (3, 128, 3, 3)
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in sd
