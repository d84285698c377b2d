"""Full-UNet oracle runs against latents produced by the reference itself (through its own
get_qmodel loader) -- tests/golden/unet_*.pt.  CPU; the SDXL / grouped cases take minutes and run
only with DGQ_SLOW=1 (they were run when the fixtures were minted; result recorded in DESIGN.md)."""
import os

import pytest
import torch

from tests import unet_cases as U

SLOW = os.environ.get("DGQ_SLOW") == "1"


@pytest.mark.parametrize("model_type,case", [
    ("sd", "w8a8_g1"),
    pytest.param("sd", "w4a8_g8_log", marks=pytest.mark.skipif(not SLOW, reason="minutes on CPU; DGQ_SLOW=1")),
    pytest.param("sdxl", "w4a8_g16_ta", marks=pytest.mark.skipif(not SLOW, reason="minutes on CPU; DGQ_SLOW=1")),
    pytest.param("sdxl", "w8a6_g1", marks=pytest.mark.skipif(not SLOW, reason="minutes on CPU; DGQ_SLOW=1")),
])
def test_oracle_matches_reference_latents(model_type, case):
    gold = U.load_golden(model_type, case)
    outs, _, _, _ = U.oracle_outputs(model_type, case)
    for k, (y, g) in enumerate(zip(outs, gold["outs"])):
        err = ((y - g).abs().max() / g.abs().max()).item()
        assert err < 2e-4, (k, err)        # same fp32 arithmetic; differences = summation order only
        assert U.cosine(y, g) > 0.999999
