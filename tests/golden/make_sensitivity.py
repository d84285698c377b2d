"""Self-sensitivity of the reference algorithm on the full-UNet parity cases.

A random-init quantized UNet is a chaotic map: a relative perturbation eps of every QuantLayer
output flips quantizer codes downstream and the flips compound block by block.  This script measures
how far the ORACLE (pinned to the reference, tests/test_oracle_golden.py) moves from itself under
eps = 1e-6 (fp32 summation-order noise), 1e-4 and 2^-11 (fp16 rounding of one operand), and stores
cosine / rel-l2 of the final output in tests/golden/self_sensitivity.json.  tests/test_unet_gpu.py
uses the table to put the engine's end-to-end deviation next to the reference's own.

    python tests/golden/make_sensitivity.py            # ~10 min on 8 cores
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dgq_oracle as O, synth as S  # noqa: E402
from tests import unet_cases as U  # noqa: E402

EPS = [1e-6, 1e-4, 2.0 ** -11]


def main():
    out_path = os.path.join(ROOT, "tests", "golden", "self_sensitivity.json")
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    orig = O.quant_layer
    for model_type, cases in U.UNET_RUNS.items():
        for case in cases:
            key = f"{model_type}/{case}"
            if key in table:
                continue
            sd, cfg, acts = U.build_case(S, O, model_type, case, torch)
            O.update_group_convs(cfg, acts[0], sd)
            inp = U.case_inputs(model_type, case, 0)
            with torch.no_grad():
                y0 = O.unet_forward(model_type, sd, acts[0], cfg, *inp)
            row = {}
            for eps in EPS:
                g = torch.Generator().manual_seed(1)

                def noisy(x, sd_, act, name, cfg_, **kw):
                    y = orig(x, sd_, act, name, cfg_, **kw)
                    return y * (1 + eps * torch.randn(y.shape, generator=g))
                O.quant_layer = noisy
                try:
                    with torch.no_grad():
                        y1 = O.unet_forward(model_type, sd, acts[0], cfg, *inp)
                finally:
                    O.quant_layer = orig
                row[f"{eps:.3g}"] = {"cosine": U.cosine(y0, y1), "rel_l2": ((y0 - y1).norm() / y0.norm()).item()}
                print(key, eps, row[f"{eps:.3g}"], flush=True)
            table[key] = row
            json.dump(table, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
