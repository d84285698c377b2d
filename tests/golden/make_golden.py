"""Mint golden fixtures by EXECUTING THE REFERENCE (ugonfor/DGQ) on CPU.

Run once in the build container (needs /root/reference; the GPU box has no
copy):

    python tests/golden/make_golden.py ops          # op-level vectors  -> ops.pt
    python tests/golden/make_golden.py unet sd      # full-UNet latents -> unet_sd_*.pt
    python tests/golden/make_golden.py unet sdxl [case]   # one case per process keeps peak RSS < 60 GB
    python tests/golden/make_golden.py vae          # AutoencoderKL.decode + image postprocess -> vae.pt

Inputs are regenerated from seeds by the tests (oracle/synth.py uses numpy
PCG64 / torch CPU generators, both machine-independent), so only the
reference's OUTPUTS are stored.  The import shims are the three of SURVEY.md
Appendix A (hub stub, matplotlib stub, .cuda() identity on CPU).
"""
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference(model_type: str):
    import huggingface_hub
    huggingface_hub.cached_download = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("offline"))
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    # the repo root also holds drop-in `quant/` and `diffusers_rewrite/` packages: the
    # reference must win here, the oracle is imported through its own package
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
    sys.path[:0] = [REF, REF + "/diffusers/src", REF + "/src"]
    os.environ["DIFFUSERS_REWRITE"] = model_type
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _to = torch.nn.Module.to

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return _to(self, *a, **k)
    torch.nn.Module.to = to
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle_pkg", REPO + "/oracle/__init__.py",
                                                  submodule_search_locations=[REPO + "/oracle"])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["oracle_pkg"] = pkg
    spec.loader.exec_module(pkg)
    import importlib
    O = importlib.import_module("oracle_pkg.dgq_oracle")
    S = importlib.import_module("oracle_pkg.synth")
    return torch, O, S


# --------------------------------------------------------------------------- #
def make_ops():
    torch, O, S = import_reference("sd")
    import torch.nn as nn
    from quant.quant_layer import UniformAffineQuantizer, QuantLayer, Scaler
    from quant.quant_layer_text import T2ILogQuantizer
    from quant.adaptive_rounding import AdaRoundQuantizer, RMODE
    import diffusers_rewrite
    gold = {}

    def uaq(bits, delta, zp):
        q = UniformAffineQuantizer(bits=bits, channel_wise=False, scaler=Scaler.MINMAX)
        q.delta, q.zero_point, q.init = delta, zp, True
        return q

    def group_params(g, n, level, view):
        lab = torch.randint(0, 8, (n,), generator=g)
        lo = -(torch.rand(8, generator=g) * 3 + 1)
        hi = torch.rand(8, generator=g) * 3 + 1
        lo[0], hi[0] = 0.5, 2.0     # zp < 0
        lo[1], hi[1] = -3.0, -0.4   # zp > level-1
        d = (hi - lo) / (level - 1)
        z = torch.round(-lo / d)
        return d[lab].view(view), z[lab].view(view)

    # ---- UniformAffineQuantizer: scalar / (1,1,X) / (1,X,1), A8 and A6
    for bits in (8, 6):
        g = torch.Generator().manual_seed(100 + bits)
        x = torch.randn(2, 48, 40, generator=g) * 2
        level = 2 ** bits
        cases = {"scalar": (torch.tensor(0.0213), torch.tensor(117.0)),
                 "in": group_params(g, 40, level, (1, 1, -1)),
                 "out": group_params(g, 48, level, (1, -1, 1))}
        for k, (d, z) in cases.items():
            gold[f"uaq_a{bits}_{k}"] = dict(seed=100 + bits, delta=d, zp=z, out=uaq(bits, d, z)(x))

    # ---- T2ILogQuantizer: static delta and real-time
    g = torch.Generator().manual_seed(7)
    p = torch.softmax(torch.randn(2, 4, 33, 77, generator=g) * 3, dim=-1)
    for rt in (False, True):
        q = T2ILogQuantizer(bits=8, real_time=rt)
        if not rt:
            q.delta, q.init = torch.tensor(0.37), True
        gold[f"t2i_log_rt{int(rt)}"] = dict(seed=7, delta=torch.tensor(0.37), out=q(p))

    # ---- weight quantizer init + AdaRound hard codes
    g = torch.Generator().manual_seed(11)
    w = torch.randn(24, 16, 3, 3, generator=g) * 0.05
    for bits in (4, 8):
        wq = UniformAffineQuantizer(bits=bits, channel_wise=True, scaler=Scaler.MINMAX)
        wdq = wq(w)
        ada = AdaRoundQuantizer(wq, rmode=RMODE.LEARNED_HARD_SIGMOID, w=w)
        alpha = torch.randn(w.shape, generator=g)
        ada.alpha = nn.Parameter(alpha)
        gold[f"wq_w{bits}"] = dict(seed=11, delta=wq.delta, zp=wq.zero_point, out=wdq,
                                   alpha=alpha, ada_out=ada(w).detach())

    # ---- QuantLayer: conv 3x3 s1/s2, conv 1x1, linear; g=1 / K-wise / row-wise
    def run_layer(layer, x, wbits, abits, d, z, grouped):
        ql = QuantLayer(layer, {"bits": wbits, "channel_wise": True, "scaler": Scaler.MINMAX},
                        {"bits": abits, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True})
        ql.set_quant_state(True, False)
        ql(x)  # initialises wqtizer exactly as load_cali_model's first dummy forward does
        ql.aqtizer.delta, ql.aqtizer.zero_point, ql.aqtizer.init = d, z, True
        ql.use_group_num = grouped
        ql.set_quant_state(True, True)
        with torch.no_grad():
            return ql(x), ql.wqtizer.delta.detach(), ql.wqtizer.zero_point.detach()

    def conv_case(tag, ci, co, k, s, hw, bsz, wbits, seed):
        g = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        layer = nn.Conv2d(ci, co, k, s, k // 2)
        x = torch.randn(bsz, ci, hw, hw, generator=g)
        ho = (hw + 2 * (k // 2) - k) // s + 1
        L, K = ho * ho, ci * k * k
        cases = {"g1": (torch.tensor(0.031), torch.tensor(121.0), False),
                 "g1u": (torch.tensor(0.031), torch.tensor(121.0), True),
                 "kwise": group_params(g, K, 256, (1, -1, 1)) + (True,),
                 "rowwise": group_params(g, L, 256, (1, 1, -1)) + (True,)}
        for name, (d, z, grouped) in cases.items():
            out, wd, wz = run_layer(layer, x, wbits, 8, d, z, grouped)
            gold[f"{tag}_{name}"] = dict(seed=seed, weight=layer.weight.detach().clone(),
                                         bias=layer.bias.detach().clone(), delta=d, zp=z,
                                         wdelta=wd, wzp=wz, out=out, grouped=grouped,
                                         shape=(bsz, ci, hw, co, k, s, wbits))

    conv_case("conv3", 32, 48, 3, 1, 12, 2, 4, 21)
    conv_case("conv3s2", 32, 48, 3, 2, 12, 2, 8, 22)
    conv_case("conv1", 64, 32, 1, 1, 8, 2, 4, 23)

    for tag, wbits, seed in (("lin_w4", 4, 31), ("lin_w8", 8, 32)):
        g = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        layer = nn.Linear(96, 80)
        x = torch.randn(2, 50, 96, generator=g)
        cases = {"g1": (torch.tensor(0.027), torch.tensor(130.0)),
                 "kwise": group_params(g, 96, 256, (1, 1, -1)),
                 "rowwise": group_params(g, 50, 256, (1, -1, 1))}
        for name, (d, z) in cases.items():
            out, wd, wz = run_layer(layer, x, wbits, 8, d, z, True)
            gold[f"{tag}_{name}"] = dict(seed=seed, weight=layer.weight.detach().clone(),
                                         bias=layer.bias.detach().clone(), delta=d, zp=z,
                                         wdelta=wd, wzp=wz, out=out)

    # ---- config 1 at full size: strided subsample of the output
    g = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    layer = nn.Conv2d(320, 320, 3, 1, 1)
    x = torch.randn(1, 320, 64, 64, generator=g)
    for name, (d, z, grouped) in {
            "g1": (torch.tensor(0.035), torch.tensor(128.0), False),
            "kwise": group_params(g, 2880, 256, (1, -1, 1)) + (True,),
            "rowwise": group_params(g, 4096, 256, (1, 1, -1)) + (True,)}.items():
        out, wd, wz = run_layer(layer, x, 4, 8, d, z, grouped)
        gold[f"config1_{name}"] = dict(seed=0, delta=d, zp=z, wdelta=wd, wzp=wz,
                                       out_sub=out.flatten()[::37].clone(), grouped=grouped)

    # ---- Attention_forward: self/cross x uniform/log(real-time) x start-peak, per-D / per-T scales
    from quant.quant_block import QuantBasicTransformerBlock
    for tag, log, rt, sp in (("uni", False, False, False), ("log_rt", True, True, False),
                             ("log_rt_sp", True, True, True), ("log_static", True, False, False)):
        seed = 40
        g = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        blk = diffusers_rewrite.BasicTransformerBlock(64)        # sd: 8 heads, d=8, ctx 768
        aq = {"bits": 8, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True}
        qb = QuantBasicTransformerBlock(blk, aq, {"softmax_a_bit": 8, "t2i_log_quant": log,
                                                  "t2i_real_time": rt, "t2i_start_peak": sp,
                                                  "log_max_1": False})
        x = torch.randn(2, 36, 64, generator=g)
        ctx = torch.randn(2, 77, 768, generator=g)
        state = {k: v.detach().clone() for k, v in qb.state_dict().items()}
        params = {}
        for an, a, S_ in (("attn1", qb.attn1, 36), ("attn2", qb.attn2, 77)):
            a.use_aq = True
            for qn, view, n in (("aqtizer_q", (1, 1, -1), 8), ("aqtizer_k", (1, -1, 1), S_ - (1 if (sp and an == "attn2") else 0)),
                                ("aqtizer_v", (1, 1, -1), 8)):
                d, z = group_params(g, n, 256, view)
                d, z = d * 0.5, z
                q = getattr(a, qn)
                q.delta, q.zero_point, q.init = d, z, True
                params[f"{an}.{qn}"] = (d, z)
            if not log:
                a.aqtizer_w.delta, a.aqtizer_w.zero_point, a.aqtizer_w.init = torch.tensor(1 / 255.), torch.tensor(0.), True
            elif not rt:
                a.aqtizer_w.delta, a.aqtizer_w.init = torch.tensor(0.41), True
        with torch.no_grad():
            o1 = qb.attn1(x)
            o2 = qb.attn2(x, encoder_hidden_states=ctx)
        gold[f"attn_{tag}"] = dict(seed=seed, state=state, params=params, out1=o1, out2=o2)

    torch.save(gold, OUT + "/ops.pt")
    print("ops.pt:", len(gold), "cases,", os.path.getsize(OUT + "/ops.pt") / 1e6, "MB")


# --------------------------------------------------------------------------- #
UNET_CASES = {
    # name: (wbits, abits, group_num, log, real_time, start_peak, n_steps, batch, timesteps)
    "w8a8_g1": (8, 8, 1, False, False, False, 1, 1, [500]),
    "w4a8_g8_log": (4, 8, 8, True, True, True, 2, 2, [981, 461]),
    "w4a8_g16_ta": (4, 8, 16, True, True, True, 1, 1, [999]),
    "w8a6_g1": (8, 6, 1, False, False, False, 1, 1, [999]),
}
UNET_RUNS = {"sd": ["w8a8_g1", "w4a8_g8_log"], "sdxl": ["w4a8_g16_ta", "w8a6_g1"]}


def build_case(S, O, model_type, case, torch):
    wb, ab, gn, log, rt, sp, n_steps, batch, ts = UNET_CASES[case]
    sd = S.make_weights(model_type, seed=0)
    S.init_weight_quant(sd, wb)
    cfg = O.QConfig(wbits=wb, abits=ab, softmax_bits=ab, t2i_log_quant=log, t2i_real_time=rt,
                    t2i_start_peak=sp)
    acts = []
    for k in range(n_steps):
        inp = S.example_inputs(model_type, batch, seed=k, t=ts[k])
        acts.append(S.calibrate_act(model_type, sd, cfg, inp, gn))
    return sd, cfg, acts


def make_unet(model_type, only=None):
    torch, O, S = import_reference(model_type)
    import diffusers_rewrite
    from quant.quant_layer import Scaler
    from quant.load_qmodel_util import get_qmodel
    import time
    for case in UNET_RUNS[model_type]:
        if only is not None and case != only:
            continue
        wb, ab, gn, log, rt, sp, n_steps, batch, ts = UNET_CASES[case]
        t0 = time.time()
        sd, cfg, acts = build_case(S, O, model_type, case, torch)
        ckpt = {"weight": sd}
        for k, a in enumerate(acts):
            ckpt[f"act_{k}"] = a
        path = f"/tmp/golden_{model_type}_{case}.pth"
        torch.save(ckpt, path)
        print(case, "synth ckpt", time.time() - t0)
        unet = diffusers_rewrite.UNet2DConditionModel()
        plain = {}
        for k, v in sd.items():
            if "wqtizer" in k:
                continue
            k2 = k[len("model."):]
            k2 = k2[:-2] + ".weight" if k2.endswith(".w") else (k2[:-2] + ".bias" if k2.endswith(".b") else k2)
            plain[k2] = v
        print(unet.load_state_dict(plain, strict=True))
        pipe = types.SimpleNamespace(unet=unet)
        # time-aware index: act_{(1000-t)//(1000//n)} (quant/calibration.py:302)
        n_inf = {1: 1, 2: 2}[n_steps]
        qnn = get_qmodel(model_type, pipe, path,
                         {"bits": wb, "channel_wise": True, "scaler": Scaler.MINMAX}, True,
                         {"bits": ab, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True},
                         {"softmax_a_bit": ab, "t2i_log_quant": log, "t2i_real_time": rt,
                          "t2i_start_peak": sp, "log_max_1": False},
                         gn > 1, n_inf, True)
        qnn.float()
        outs = []
        for k in range(n_steps):
            inp = S.example_inputs(model_type, batch, seed=k, t=ts[k])
            with torch.no_grad():
                if model_type == "sdxl":
                    y = qnn(inp[0], inp[1], inp[2], added_cond_kwargs=inp[3])[0]
                else:
                    y = qnn(inp[0], inp[1], inp[2])[0]
            outs.append(y.clone())
            print(case, "step", k, float(y.abs().mean()), time.time() - t0)
        torch.save({"outs": outs, "case": UNET_CASES[case]}, f"{OUT}/unet_{model_type}_{case}.pt")


VAE_CASES = {   # name -> (config of oracle.vae_oracle.VAE_CONFIGS, latent batch, latent size, weight seed, input seed)
    "small_b2_16": ("small", 2, 16, 0, 1),
    "sd_b1_8": ("sd", 1, 8, 0, 2),
    "sdxl_b1_8": ("sdxl", 1, 8, 3, 4),
}


def make_vae():
    """The reference's own AutoencoderKL (vendored diffusers) on weights synthesised by oracle.vae_oracle: the
    pipelines' `vae.decode(latents / scaling_factor)` and `VaeImageProcessor.postprocess`."""
    torch, O, S = import_reference("sd")
    import importlib
    V = importlib.import_module("oracle_pkg.vae_oracle")
    from diffusers.models.autoencoders.autoencoder_kl import AutoencoderKL
    from diffusers.image_processor import VaeImageProcessor
    gold = {}
    for name, (cfgname, b, size, wseed, iseed) in VAE_CASES.items():
        cfg = V.VAE_CONFIGS[cfgname]
        n = len(cfg["block_out_channels"])
        vae = AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * n,
                            up_block_types=("UpDecoderBlock2D",) * n, block_out_channels=cfg["block_out_channels"],
                            layers_per_block=cfg["layers_per_block"], latent_channels=cfg["latent_channels"],
                            norm_num_groups=32, sample_size=size * 8, scaling_factor=cfg["scaling_factor"]).eval()
        sd = V.make_vae_state(cfg, wseed)
        want = {k for k in vae.state_dict() if k.startswith("decoder.") or k.startswith("post_quant_conv.")}
        assert want == set(sd), (sorted(want - set(sd))[:5], sorted(set(sd) - want)[:5])
        print(name, vae.load_state_dict(sd, strict=False).unexpected_keys)
        g = torch.Generator().manual_seed(iseed)
        lat = torch.randn(b, cfg["latent_channels"], size, size, generator=g) * cfg["scaling_factor"] * 4.0
        with torch.no_grad():
            img = vae.decode(lat / vae.config.scaling_factor, return_dict=False)[0]
        proc = VaeImageProcessor(vae_scale_factor=2 ** (n - 1))
        np_img = proc.postprocess(img, output_type="np")
        u8 = torch.from_numpy((np_img * 255).round().astype("uint8"))          # numpy_to_pil's conversion
        gold[name] = {"case": VAE_CASES[name], "latents": lat, "image": img.clone(), "u8": u8}
        print(name, tuple(img.shape), float(img.abs().mean()), float(u8.float().mean()))
    torch.save(gold, f"{OUT}/vae.pt")


if __name__ == "__main__":
    if sys.argv[1] == "ops":
        make_ops()
    elif sys.argv[1] == "vae":
        make_vae()
    else:
        make_unet(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
