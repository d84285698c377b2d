"""Counterpart of the reference's src/inference_qmodel.py: same quantization flags, the
quantized UNet + the device sampler loop + (optionally) the VAE decode and the PNG files the reference writes
(src/inference_qmodel.py:46-53).  The text encoders stay outside: prompt embeddings come from a file or a seeded draw
(no pretrained weights exist offline).

  DIFFUSERS_REWRITE=sdxl python scripts/inference_qmodel.py --cali_ckpt ckpt_merged.pth --wq 4 --use_aq --aq 8 \
      --use_group --t2i_log_quant --t2i_real_time --t2i_start_peak --time_aware_aqtizer \
      [--num_inference_steps N] [--prompt_embeds embeds.pt] [--batch 2] [--out latents.pt]
  ... --compiled unet.dgqb          # a dgq_b200 compiled checkpoint instead of --cali_ckpt (flags come from its header)
  ... --synthetic                   # random-init weights + synthetic scales (no checkpoint at all)
  ... --vae vae_state.pt --outdir imgs      # decode with a diffusers AutoencoderKL state dict and save PNGs
  ... --synthetic_vae --outdir imgs         # the same path on random-init decoder weights

`--prompt_embeds`: a torch file {"ctx": [B,77,C]} (SD: the cond embeddings; the uncond half {"uncond": ...} is
optional, zeros otherwise) or, for SDXL, {"ctx": [B,77,2048], "text_embeds": [B,1280]}; omitted => N(0,1) draws
with --seed.  Writes the final latents (fp32, [B,4,h,w]) and, with --vae / --synthetic_vae, the decoded images as
tmp_{model_type}_{i}_{precision}.png (the reference's naming without the prompt text).
Defaults follow the reference: 25 PLMS steps + CFG 7.5 for sd, 4 Euler-ancestral steps, guidance 0 for sdxl."""
import argparse
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
MODEL_TYPE = os.environ.get("DIFFUSERS_REWRITE", "sd")


def parse_args():
    p = argparse.ArgumentParser(description="DGQ quantized UNet + sampler on dgq_b200 (latents in, latents out)")
    p.add_argument("--use_group", action="store_true")
    p.add_argument("--num_inference_steps", type=int, default=-1)
    p.add_argument("--cali_ckpt", type=str, default=None)
    p.add_argument("--compiled", type=str, default=None)
    p.add_argument("--synthetic", action="store_true")
    p.add_argument("--fp16", action="store_true", help="accepted for CLI parity: operands are fp16 either way")
    p.add_argument("--wq", type=int, default=4)
    p.add_argument("--use_aq", action="store_true")
    p.add_argument("--aq", type=int, default=8)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--t2i_log_quant", action="store_true")
    p.add_argument("--t2i_real_time", action="store_true")
    p.add_argument("--t2i_start_peak", action="store_true")
    p.add_argument("--time_aware_aqtizer", action="store_true")
    p.add_argument("--prompt_embeds", type=str, default=None)
    p.add_argument("--batch", type=int, default=2, help="images (the reference runs [prompt] * 2)")
    p.add_argument("--guidance", type=float, default=None)
    p.add_argument("--out", type=str, default=None)
    p.add_argument("--vae", type=str, default=None, help="AutoencoderKL state dict (torch.save / safetensors): decode + save PNGs")
    p.add_argument("--synthetic_vae", action="store_true", help="random-init VAE decoder weights (path check only)")
    p.add_argument("--outdir", type=str, default=".")
    return p.parse_args()


def build_vae(opt):
    """dgq_b200.vae.VaeDecoder with the stabilityai sd-vae / sdxl-vae decoder architecture"""
    import torch
    from dgq_b200.vae import VaeDecoder
    vae = VaeDecoder(scaling_factor=0.18215 if MODEL_TYPE == "sd" else 0.13025)
    if opt.vae:
        if opt.vae.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(opt.vae)
        else:
            sd = torch.load(opt.vae, map_location="cpu")
        res = vae.load_state_dict(sd, strict=False)
        if res.missing_keys:
            raise SystemExit(f"--vae: {len(res.missing_keys)} decoder tensors missing, e.g. {res.missing_keys[:3]}")
    else:
        g = torch.Generator().manual_seed(opt.seed)
        with torch.no_grad():
            for prm in vae.parameters():
                if prm.dim() > 1:
                    prm.copy_((torch.rand(prm.shape, generator=g) * 2 - 1) / (prm[0].numel() ** 0.5))
    return vae.cuda().eval()


def build_qnn(opt, n_steps):
    import torch
    from dgq_b200 import compiled, synthetic
    from quant.load_qmodel_util import get_qmodel
    from quant.quant_layer import Scaler
    if opt.compiled:
        return compiled.load_compiled(opt.compiled, "cuda")
    if opt.synthetic:
        return synthetic.make_qmodel(MODEL_TYPE, wbits=opt.wq, abits=opt.aq, group_num=16 if opt.use_group else 1,
                                     n_steps=n_steps, log_quant=opt.t2i_log_quant, real_time=opt.t2i_real_time,
                                     start_peak=opt.t2i_start_peak, device="cuda", seed=opt.seed)
    if not opt.cali_ckpt:
        raise SystemExit("one of --cali_ckpt, --compiled, --synthetic is required")
    pipe = types.SimpleNamespace(unet=synthetic.build_unet(MODEL_TYPE, "cuda"))
    wq_params = {"bits": opt.wq, "channel_wise": True, "scaler": Scaler.MINMAX}
    aq_params = {"bits": opt.aq, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": opt.use_aq}
    softmax = {"softmax_a_bit": opt.aq, "t2i_log_quant": opt.t2i_log_quant, "t2i_real_time": opt.t2i_real_time,
               "t2i_start_peak": opt.t2i_start_peak, "log_max_1": False}
    qnn = get_qmodel(MODEL_TYPE, pipe, opt.cali_ckpt, wq_params, opt.use_aq, aq_params, softmax, opt.use_group,
                     num_inference_steps=n_steps, time_aware_aqtizer=opt.time_aware_aqtizer if opt.use_aq else False)
    qnn.disable_out_quantization()
    return qnn


def main():
    opt = parse_args()
    import torch
    from dgq_b200 import sampler
    if not torch.cuda.is_available():
        raise SystemExit("dgq_b200 needs a CUDA device (no CPU fallback)")
    n_steps = opt.num_inference_steps if opt.num_inference_steps > 0 else (25 if MODEL_TYPE == "sd" else 4)
    t0 = time.time()
    qnn = build_qnn(opt, n_steps)
    qnn.enable_cuda_graphs(True)
    t_load = time.time() - t0
    g = torch.Generator().manual_seed(opt.seed)
    b = opt.batch
    emb = torch.load(opt.prompt_embeds, map_location="cpu") if opt.prompt_embeds else {}
    dev = "cuda"
    with torch.no_grad():
        if MODEL_TYPE == "sdxl":
            ctx = emb.get("ctx", torch.randn(b, 77, 2048, generator=g)).float().to(dev)
            b = ctx.shape[0]
            added = {"text_embeds": emb.get("text_embeds", torch.randn(b, 1280, generator=g)).float().to(dev),
                     "time_ids": torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]]).repeat(b, 1).to(dev)}
            lat = torch.randn(b, 4, 128, 128, generator=g).to(dev)
            noises = [torch.randn(b, 4, 128, 128, generator=g).to(dev) for _ in range(n_steps)]
            t1 = time.time()
            x = sampler.denoise_sdxl(qnn, lat, ctx, added, n_steps, noises)
        else:
            cond = emb.get("ctx", torch.randn(b, 77, 768, generator=g)).float()
            b = cond.shape[0]
            uncond = emb.get("uncond", torch.zeros_like(cond)).float()
            guidance = 7.5 if opt.guidance is None else opt.guidance
            ctx = (torch.cat([uncond, cond]) if guidance > 1.0 else cond).to(dev)
            lat = torch.randn(b, 4, 64, 64, generator=g).to(dev)
            t1 = time.time()
            x = sampler.denoise_sd(qnn, lat, ctx, n_steps, guidance=guidance)
        torch.cuda.synchronize()
    t_run = time.time() - t1
    out = opt.out or f"latents_{MODEL_TYPE}_w{opt.wq}a{opt.aq if opt.use_aq else 32}_{n_steps}steps.pt"
    torch.save({"latents": x.float().cpu(), "model_type": MODEL_TYPE, "steps": n_steps, "seed": opt.seed}, out)
    if opt.vae or opt.synthetic_vae:
        from dgq_b200 import vae as vae_mod
        vae = build_vae(opt)
        t2 = time.time()
        img = vae.decode_latents(x.float())             # pipeline_stable_diffusion.py:1066-1069 / _xl.py:1295-1307
        u8 = vae_mod.postprocess(img)
        torch.cuda.synchronize()
        precision = f"w{opt.wq}a{opt.aq if opt.use_aq else 32}{'g?' if opt.use_group else 'g1'}"
        paths = vae_mod.save_images(u8, opt.outdir, prefix=f"tmp_{MODEL_TYPE}_{precision}")
        print(f"vae decode {tuple(img.shape)} {time.time() - t2:.2f} s -> {paths[0]} ... ({len(paths)} files)")
    print(f"{MODEL_TYPE}: load {t_load:.1f} s, {n_steps}-step loop for {b} images {t_run:.2f} s (incl. graph capture), "
          f"latents {tuple(x.shape)} finite={bool(torch.isfinite(x).all())} -> {out}")


if __name__ == "__main__":
    main()
