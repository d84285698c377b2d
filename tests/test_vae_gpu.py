"""VAE decode on the GPU: dgq_b200.vae.VaeDecoder through the C ABI against
(a) outputs of the reference's own AutoencoderKL (tests/golden/vae.pt) and (b) the CPU oracle at a larger size.
The decoder is full precision in the reference; here operands are fp16 (activations and weights rounded to 11 bits,
fp32 accumulation), so the bar is a floating-point one: max error <= 1e-2 of the output range, cosine >= 0.9999,
8-bit images within one level on >= 99 % of the pixels."""
import math
import os

import pytest
import torch

from oracle import vae_oracle as V

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vae.pt")
MAX_ERR, COS_BAR, U8_WITHIN_1 = 1e-2, 0.9999, 0.99


def build(cfgname, wseed):
    from dgq_b200.vae import VaeDecoder
    cfg = V.VAE_CONFIGS[cfgname]
    sd = V.make_vae_state(cfg, wseed)
    vae = VaeDecoder(cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"], cfg["scaling_factor"])
    missing = vae.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return vae.cuda().eval(), sd, cfg


def check(img, ref, u8=None, ref_u8=None):
    err = float((img - ref).abs().max() / ref.abs().max())
    cos = float(torch.nn.functional.cosine_similarity(img.flatten(), ref.flatten(), dim=0))
    assert err <= MAX_ERR and cos >= COS_BAR, (err, cos)
    if u8 is not None:
        d = (u8.int() - ref_u8.int()).abs()
        assert float((d <= 1).float().mean()) >= U8_WITHIN_1 and int(d.max()) <= 4, (float((d <= 1).float().mean()), int(d.max()))
    return err, cos


@pytest.mark.parametrize("name", ["small_b2_16", "sd_b1_8", "sdxl_b1_8"])
def test_decode_vs_reference_golden(name):
    from dgq_b200 import ops, vae as vae_mod
    g = torch.load(GOLD, weights_only=False)[name]
    cfgname, b, size, wseed, iseed = g["case"]
    vae, sd, cfg = build(cfgname, wseed)
    n0 = ops.LAUNCHES
    img = vae.decode_latents(g["latents"].cuda())
    assert ops.LAUNCHES > n0 and img.shape == g["image"].shape and img.is_cuda
    u8 = vae_mod.postprocess(img)
    err, cos = check(img.cpu(), g["image"], u8, g["u8"])
    print(f"[vae] {name}: max err / range {err:.2e}, cosine {cos:.6f}")


def test_decode_vs_oracle_larger_and_chunked():
    """sd decoder, 2 latents of 32 x 32 (attention over 1024 tokens, 256 x 256 images); a chunk limit that forces
    one image per pass gives the identical result"""
    vae, sd, cfg = build("sd", 5)
    lat = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(7)) * cfg["scaling_factor"] * 3.0
    with torch.no_grad():
        ref = V.decode_latents(sd, cfg, lat)
    img = vae.decode_latents(lat.cuda())
    err, cos = check(img.cpu(), ref)
    print(f"[vae] sd 2x32x32: max err / range {err:.2e}, cosine {cos:.6f}")
    vae.CHUNK_BYTES = 1
    img1 = vae.decode_latents(lat.cuda())
    check(img1.cpu(), ref)
    assert float((img1 - img).abs().max()) <= 1e-3 * float(ref.abs().max())   # GroupNorm statistics are per sample


def test_softmax_rows():
    from dgq_b200 import ops
    g = torch.Generator().manual_seed(0)
    for rows, cols in ((64, 1024), (7, 4100), (300, 16384)):
        s = (torch.randn(rows, cols, generator=g) * 30).cuda()
        scale = 1.0 / math.sqrt(512)
        p = ops.softmax_rows(s, scale)
        ref = torch.softmax(s * scale, dim=-1)
        assert p.dtype == torch.float16 and p.shape == s.shape
        assert float((p.float() - ref).abs().max()) <= 2e-3 * float(ref.max())
        assert float((p.float().sum(-1) - 1).abs().max()) < 2e-3


def test_cpu_tensor_raises():
    from dgq_b200.vae import VaeDecoder
    vae = VaeDecoder((128, 128, 256, 256), 1)
    with pytest.raises(RuntimeError):
        vae.decode(torch.zeros(1, 4, 8, 8))
