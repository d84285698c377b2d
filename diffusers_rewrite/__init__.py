"""Drop-in `diffusers_rewrite` package: SD-v1.4 or SDXL UNet graph selected by the environment
variable DIFFUSERS_REWRITE ("sd" default, "sdxl"), read once at import like the reference
(diffusers_rewrite/__init__.py:1-6)."""
import os
import sys

from dgq_b200.unet import sd, sdxl  # noqa: F401

sys.modules[__name__ + ".sd"] = sd
sys.modules[__name__ + ".sdxl"] = sdxl
if os.environ.get("DIFFUSERS_REWRITE", "sd") == "sdxl":
    from dgq_b200.unet.sdxl import *  # noqa: F401,F403
else:
    from dgq_b200.unet.sd import *  # noqa: F401,F403
