"""Full sampler-loop parity on the GPU box: SD W4A8 g8 + t2i-log (real-time, start-peak), time-aware scales,
PLMS with classifier-free guidance -- dgq_b200 (QuantModel + device sampler) vs the CPU oracle (oracle UNet +
oracle sampler) on the same synthetic checkpoint and inputs.  BASELINE.json configs[2] at a reduced step count:

    python scripts/sd_loop_parity.py --steps 10 [--latents 1]

Prints the final-latent cosine; writes gpurun_out/sd_loop_parity.json."""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from oracle import dgq_oracle as O, sampler_oracle as SO, synth as S  # noqa: E402
from tests import unet_cases as U  # noqa: E402
from tests.golden import make_golden as MG  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--latents", type=int, default=1)
    ap.add_argument("--guidance", type=float, default=7.5)
    a = ap.parse_args()
    from dgq_b200 import sampler as DS
    torch.set_num_threads(os.cpu_count() or 1)
    model_type, case = "sd", "w4a8_g8_log"
    sd, cfg, acts2 = U.build_case(S, O, model_type, case, torch)
    acts = [acts2[k % len(acts2)] for k in range(a.steps)]       # one table per step index
    MG.UNET_CASES["loop"] = MG.UNET_CASES[case][:6] + (a.steps, 2 * a.latents, [0] * a.steps)
    with tempfile.TemporaryDirectory() as tmp:
        qnn = U.build_qmodel(model_type, "loop", sd, acts, tmp)
    qnn.enable_cuda_graphs(True)
    g = torch.Generator().manual_seed(7)
    lat = torch.randn(a.latents, 4, 64, 64, generator=g)
    ctx = torch.randn(2 * a.latents, 77, 768, generator=g)

    def oracle_unet(x, t, c):
        idx = int((1000 - float(t)) // (1000 // a.steps))
        O.update_group_convs(cfg, acts[idx], sd)
        return O.unet_forward(model_type, sd, acts[idx], cfg, x, t, c)
    with torch.no_grad():
        t0 = time.time()
        got = DS.denoise_sd(qnn, lat.cuda(), ctx.cuda(), a.steps, guidance=a.guidance)
        torch.cuda.synchronize()
        t_gpu = time.time() - t0
        t0 = time.time()
        want = SO.denoise_sd(oracle_unet, lat, ctx, a.steps, guidance=a.guidance)
        t_cpu = time.time() - t0
    cos = U.cosine(got, want)
    l2 = ((got.cpu() - want).norm() / want.norm()).item()
    res = {"steps": a.steps, "unet_calls": a.steps + 1, "latents": a.latents, "guidance": a.guidance, "cosine": cos,
           "rel_l2": l2, "gpu_s_incl_graph_capture": round(t_gpu, 2), "cpu_oracle_s": round(t_cpu, 1),
           "cores": os.cpu_count()}
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/sd_loop_parity.json", "w"))


if __name__ == "__main__":
    main()
