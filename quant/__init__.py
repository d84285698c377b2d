"""Drop-in `quant` package: the import paths of ugonfor/DGQ (`quant.quant_model.QuantModel`,
`quant.quant_layer.QuantLayer`, `quant.load_qmodel_util.get_qmodel`, ...) bound to dgq_b200's
CUDA-backed implementation.  The sub-modules are aliases, so class identity holds across both
spellings (the vendored SDXL pipeline checks `type(unet) == QuantModel`)."""
import importlib
import sys

for _name in ("quant_layer", "quant_layer_text", "adaptive_rounding", "quant_block", "quant_model",
              "calibration", "load_qmodel_util"):
    _mod = importlib.import_module(f"dgq_b200.quant.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
