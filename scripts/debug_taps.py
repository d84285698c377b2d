"""Block-by-block deviation of the CUDA engine from the CPU oracle on one UNet case."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import dgq_oracle as O, synth as S
from tests import unet_cases as U
from dgq_b200 import engine

model_type, case = sys.argv[1], sys.argv[2]
sd, cfg, acts = U.build_case(S, O, model_type, case, torch)
O.update_group_convs(cfg, acts[0], sd)
inp = U.case_inputs(model_type, case, 0)
taps = []
with torch.no_grad():
    y_ref = O.unet_forward(model_type, sd, acts[0], cfg, *inp, taps=taps)
qnn = U.build_qmodel(model_type, case, sd, acts, tempfile.mkdtemp())
engine.TAPS = []
y = U.run_qmodel(qnn, model_type, case, 0)
for (n1, a), (n2, b) in zip(taps, engine.TAPS):
    b = b.cpu()
    if b.dim() == 2 and a.dim() == 2:
        b = b[:, : a.shape[1]]
    print(f"{n1:14s} {n2:14s} rel-l2 {((a - b).norm() / a.norm()).item():.5f}  cos {U.cosine(a, b):.6f}")
print("final", ((y.cpu() - y_ref).norm() / y_ref.norm()).item(), U.cosine(y, y_ref))
