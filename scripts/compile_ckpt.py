"""One-shot converter: reference `*_merged .pth` -> dgq_b200 compiled checkpoint.

  python scripts/compile_ckpt.py --model sd|sdxl --ckpt cali_ckpt_merged.pth --out unet.dgqb \
      --wbits 4 --abits 8 --steps 50 [--group] [--time-aware] [--no-log-quant] [--no-real-time] [--no-start-peak]

Loads the checkpoint exactly as the reference's inference entry point does (quant.load_qmodel_util.get_qmodel
around a random-init UNet skeleton: every weight comes from the checkpoint's 'weight' dict), then writes the
packed file.  Needs a CUDA device (weight packing runs on the GPU).  Load it back with
dgq_b200.compiled.load_compiled(path)."""
import argparse
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", choices=["sd", "sdxl"], required=True)
    ap.add_argument("--ckpt", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--wbits", type=int, default=4)
    ap.add_argument("--abits", type=int, default=8)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--group", action="store_true")
    ap.add_argument("--time-aware", action="store_true")
    ap.add_argument("--no-log-quant", action="store_true")
    ap.add_argument("--no-real-time", action="store_true")
    ap.add_argument("--no-start-peak", action="store_true")
    a = ap.parse_args()
    os.environ["DIFFUSERS_REWRITE"] = a.model
    import torch
    from dgq_b200 import compiled, synthetic
    from quant.load_qmodel_util import get_qmodel
    from quant.quant_layer import Scaler
    t0 = time.time()
    pipe = types.SimpleNamespace(unet=synthetic.build_unet(a.model, "cuda"))
    qnn = get_qmodel(a.model, pipe, a.ckpt, {"bits": a.wbits, "channel_wise": True, "scaler": Scaler.MINMAX}, True,
                     {"bits": a.abits, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True},
                     {"softmax_a_bit": a.abits, "t2i_log_quant": not a.no_log_quant, "t2i_real_time": not a.no_real_time,
                      "t2i_start_peak": not a.no_start_peak, "log_max_1": False},
                     a.group, a.steps, a.time_aware)
    t1 = time.time()
    h = compiled.compile_checkpoint(qnn, a.out)
    torch.cuda.synchronize()
    print(f"loaded {a.ckpt} in {t1 - t0:.1f} s, wrote {a.out}: {os.path.getsize(a.out) / 2**20:.0f} MiB, "
          f"{len(h['layers'])} QuantLayers, {len(h['act'])} activation tables, sha256 {h['sha256'][:16]}... "
          f"in {time.time() - t1:.1f} s")


if __name__ == "__main__":
    main()
