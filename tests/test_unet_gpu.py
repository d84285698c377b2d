"""End-to-end parity of the CUDA engine behind the reference's own API
(UNet2DConditionModel -> get_qmodel(ckpt) -> qnn(sample, t, ctx)) against latents the REFERENCE
produced on the same synthetic checkpoint and inputs (tests/golden/unet_*.pt).

north_star bar: final-latent cosine >= 0.999.  It is asserted AS IS on every case where the reference can meet
it against ITSELF.  A random-init quantized UNet is a chaotic map (one flipped 8-bit code changes the next
quantizer's input, flips compound block by block); tests/golden/self_sensitivity.json records how far the
reference moves from its own output when every QuantLayer result is perturbed by 1e-6 relative -- fp32
summation-order noise, i.e. what a different BLAS already does:

    case                      reference vs itself (1e-6)    bar asserted here
    sd/w8a8_g1                0.99898                       WAIVED to self - 0.002 (the reference misses 0.999 itself)
    sd/w4a8_g8_log            0.99936                       0.999
    sdxl/w4a8_g16_ta (headline) 0.99972                     0.999
    sdxl/w8a6_g1              0.99024                       WAIVED to self - 0.002 (six-bit codes: 4x coarser flips)

The two waivers are stated here, in DESIGN.md section 5 and in the test output; they are not a tolerance on the
kernels -- per-layer parity is asserted without any waiver by tests/test_layerwise_gpu.py (teacher-forced: every
layer of these same four cases fed the oracle's input, codes bit-exact, outputs <= 1e-2)."""
import json
import os

import pytest
import torch

NORTH_STAR = 0.999
_SENS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "self_sensitivity.json")))


def cos_bar(model_type, case):
    """(bar, waived): 0.999 wherever the reference reaches it against itself under 1e-6 noise."""
    self_cos = _SENS[f"{model_type}/{case}"]["1e-06"]["cosine"]
    if self_cos >= NORTH_STAR + 2e-4:
        return NORTH_STAR, False
    return self_cos - 0.002, True


from oracle import dgq_oracle as O, synth as S  # noqa: E402
from tests import unet_cases as U  # noqa: E402

pytestmark = pytest.mark.gpu


def test_bars_are_the_documented_ones():
    assert cos_bar("sd", "w4a8_g8_log") == (NORTH_STAR, False)
    assert cos_bar("sdxl", "w4a8_g16_ta") == (NORTH_STAR, False)
    assert cos_bar("sd", "w8a8_g1")[1] and cos_bar("sdxl", "w8a6_g1")[1]


@pytest.mark.parametrize("model_type,case", [("sd", "w8a8_g1"), ("sd", "w4a8_g8_log"),
                                             ("sdxl", "w4a8_g16_ta"), ("sdxl", "w8a6_g1")])
def test_unet_matches_reference(model_type, case, tmp_path):
    from dgq_b200 import ops
    gold = U.load_golden(model_type, case)
    sd, cfg, acts = U.build_case(S, O, model_type, case, torch)
    qnn = U.build_qmodel(model_type, case, sd, acts, tmp_path)
    bar, waived = cos_bar(model_type, case)
    n0 = ops.LAUNCHES
    report = []
    for k, g in enumerate(gold["outs"]):
        y = U.run_qmodel(qnn, model_type, case, k)
        assert y.shape == g.shape and y.dtype == g.dtype
        assert torch.isfinite(y).all()
        cos = U.cosine(y, g)
        l2 = ((y.cpu() - g).norm() / g.norm()).item()
        report.append(dict(step=k, cosine=cos, rel_l2=l2))
        note = (f"north_star 0.999 WAIVED: the reference reaches {bar + 0.002:.5f} against itself under 1e-6 noise"
                if waived else "north_star 0.999 asserted")
        print(f"{model_type}/{case} step {k}: cosine {cos:.6f} rel-l2 {l2:.4f} (bar {bar:.5f}; {note})")
        assert cos >= bar, (k, cos, bar)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(dict(case=f"{model_type}/{case}", bar=bar, waived=waived, steps=report),
              open(f"gpurun_out/unet_cosine_{model_type}_{case}.json", "w"), indent=1)
    assert ops.LAUNCHES > n0  # the CUDA kernels ran (no eager fallback exists)
    del qnn
    torch.cuda.empty_cache()
