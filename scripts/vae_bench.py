"""Time dgq_b200.vae.VaeDecoder.decode on the two pipelines' sizes (SD 64x64 latents -> 512x512, SDXL 128x128 ->
1024x1024), random-init decoder weights, CUDA events after warm-up; prints one JSON line per case.

    python scripts/vae_bench.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dgq_b200 import ops  # noqa: E402
from dgq_b200.vae import VaeDecoder  # noqa: E402


def decoder_gmac(h, w, boc=(128, 256, 512, 512), lpb=2):
    """multiply-accumulates of one decode of an h x w latent (convs + the mid-block attention), in GMAC"""
    rev = list(reversed(boc))
    mac = h * w * (4 * 4 + 9 * 4 * rev[0])
    c = rev[0]
    mac += 4 * h * w * 9 * c * c + 4 * h * w * c * c + 2 * (h * w) ** 2 * c     # mid: 2 resnets, q k v o, q k^T and p v
    cout = rev[0]
    for i, cc in enumerate(rev):
        cin, cout = cout, cc
        for j in range(lpb + 1):
            ci = cin if j == 0 else cout
            mac += h * w * (9 * ci * cout + 9 * cout * cout + (ci * cout if ci != cout else 0))
        if i != len(rev) - 1:
            h, w = 2 * h, 2 * w
            mac += h * w * 9 * cout * cout
    mac += h * w * 9 * boc[0] * 3
    return mac / 1e9


def main():
    dev = "cuda"
    out = []
    for name, size, sf, batches in (("sd", 64, 0.18215, (1, 8)), ("sdxl", 128, 0.13025, (1, 4))):
        vae = VaeDecoder(scaling_factor=sf).to(dev).eval()
        for b in batches:
            lat = torch.randn(b, 4, size, size, device=dev) * sf
            for _ in range(2):
                vae.decode_latents(lat)
            n0 = ops.LAUNCHES
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            reps = 3
            for _ in range(reps):
                vae.decode_latents(lat)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            g = decoder_gmac(size, size) * b
            row = {"case": f"{name} vae decode {size}x{size} latents -> {size * 8}x{size * 8}, batch {b}", "ms": round(ms, 2),
                   "images_per_s": round(b / ms * 1e3, 2), "tflops": round(2 * g / ms, 1),
                   "launches": (ops.LAUNCHES - n0) // reps, "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
            print(json.dumps(row), flush=True)
            out.append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/vae_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
