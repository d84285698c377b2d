#!/usr/bin/env python
"""bench.py -- DGQ quantized-UNet forward path on B200: one JSON line per run.

  python bench.py [--config 1..5] [--gpus N] [--steps K] [--warmup W]     # dgq_b200 (CUDA) arm
  python bench.py --impl reference [--config ..] [--steps K] [--warmup W] # the reference's CPU fake-quant path
  torchrun --nproc-per-node N ... bench.py --gpus N ...                   # one rank per GPU

--config selects one of BASELINE.json's configs (default 4, the one `metric` is quoted on):
  1  single QuantLayer Conv2d 320->320 3x3 on 1x320x64x64, W4A8 group 8 (K-wise group scales)   [step = one forward]
  2  SD v1.4 UNet, one denoising step, W8A8 no grouping, batch 1                                [step = one UNet call]
  3  SD v1.4 W4A8 group 8 + t2i-log (real-time, start-peak), 50-step PLMS, 8 prompts + CFG       [step = the 51-call loop]
  4  SDXL-turbo W4A8 group 16 time-aware, 1 step, 128x128 latent, batch 16 per GPU (weak scaling) [step = one UNet call]
  5  SDXL W8A6 group 1, prompt-batch sweep 8 -> 512 sharded over the N GPUs (strong scaling)      [step = 512 images]

  value        : inputs resident in HBM, CUDA-graph replay, CUDA events, max over ranks
  e2e          : the same work through QuantModel.__call__ with HOST (pinned) inputs -- H2D of latents / prompt
                 embeddings / conditioning and D2H of the result inside the timed region
  roofline     : the qGEMM kernels (dominant): QuantLayer FLOPs / summed qGEMM time (per-launch CUDA events)
  roofline_attention, tail : the attention kernel against the tensor peak, the HBM-bound kernels as GB/s
  cpu_baseline : oracle port of the reference's fake-quant forward on the host cores, bounded sample
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line

# QuantLayer GEMM / attention MACs per sample per UNet call
GMAC = {"sd": (338.61, 63.03), "sdxl": (2988.66, 391.96)}
CONFIGS = {
    1: dict(model="layer", wbits=4, abits=8, groups=8, batch=1, metric="QuantLayer Conv2d 320->320 3x3 W4A8 g8 forwards/sec",
            unit="forwards/s", workload="single QuantLayer Conv2d 320->320 3x3 on 1x320x64x64, W4A8 group=8 K-wise scales (BASELINE configs[0])"),
    2: dict(model="sd", wbits=8, abits=8, groups=1, batch=1, n_steps=1, log=False, metric="SD v1.4 W8A8 UNet images/sec (batch 1)",
            unit="images/s", workload="sd-v1.4 W8A8 no grouping, uniform softmax quant, 1 step, 64x64 latent, batch 1 (BASELINE configs[1])"),
    3: dict(model="sd", wbits=4, abits=8, groups=8, batch=16, n_steps=50, log=True, prompts=8,
            metric="SD v1.4 W4A8 g8 50-step PLMS images/sec", unit="images/s",
            workload="sd-v1.4 W4A8 g8 time-aware t2i-log(real-time,start-peak), 50-step PLMS (51 UNet calls), 8 prompts + CFG = UNet batch 16 (BASELINE configs[2])"),
    4: dict(model="sdxl", wbits=4, abits=8, groups=16, batch=16, n_steps=1, log=True, metric="SDXL-turbo W4A8 UNet images/sec",
            unit="images/s", workload="sdxl-turbo W4A8 g16 time-aware t2i-log(real-time,start-peak), 1 step, 128x128 latent, batch 16/GPU"),
    5: dict(model="sdxl", wbits=8, abits=6, groups=1, batch=16, n_steps=1, log=False, total=512,
            metric="SDXL-turbo W8A6 g1 UNet images/sec (512-prompt batch)", unit="images/s",
            workload="sdxl-turbo W8A6 group=1 uniform softmax quant, 1 step, 128x128 latent, prompt batch 8->512 sharded over the GPUs, micro-batch <= 16 (BASELINE configs[4])"),
}


def peaks():
    try:
        pk, src = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        pk, src = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"
    try:   # int8 tensor peak measured the same way (scripts/measure_i8_peak.py -> profiles/r2_i8_peak.json)
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_i8_peak.json")))
        pk["int8_tops"], pk["int8_tops_sustained"] = d["int8_tops"], d["int8_tops_sustained"]
    except Exception:
        pass
    return pk, src


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_inputs(torch, cfg, batch, seed, device="cpu", pin=False):
    """[sample, t, ctx(, text_embeds, time_ids)] of SURVEY.md 8d for one UNet call of `batch` samples."""
    g = torch.Generator().manual_seed(seed)
    if cfg["model"] == "sdxl":
        sigma = 14.6146
        ts = [torch.randn(batch, 4, 128, 128, generator=g) * sigma / (sigma ** 2 + 1) ** 0.5, torch.tensor([999.0]),
              torch.randn(batch, 77, 2048, generator=g), torch.randn(batch, 1280, generator=g),
              torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]]).repeat(batch, 1)]
    else:
        ts = [torch.randn(batch, 4, 64, 64, generator=g), torch.tensor([500.0 if cfg.get("n_steps", 1) == 1 else 981.0]),
              torch.randn(batch, 77, 768, generator=g)]
    if pin:
        ts = [x.pin_memory() for x in ts]
    return [x.to(device) for x in ts]


def call_unet(qnn, inp):
    if len(inp) == 5:
        return qnn(inp[0], inp[1], inp[2], added_cond_kwargs={"text_embeds": inp[3], "time_ids": inp[4]})[0]
    return qnn(inp[0], inp[1], inp[2])[0]


# ----------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from dgq_b200 import ops, synthetic, engine

    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if not os.path.exists(os.path.join(ROOT, "dgq_b200", "_C", "libdgq_b200.so")):
        raise SystemExit("libdgq_b200.so missing: run __graft_entry__.build() (no fallback path)")
    warmup = max(args.warmup, 3)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    extra, breakdown, qnn = {}, None, None
    if cfg["model"] == "layer":
        step_resident, step_e2e, launches_per_step, units, h2d, d2h, bd_fn = setup_layer(torch, cfg, dev)
        scaling = "weak"
    else:
        # every rank holds the SAME replica (seed 0): a replicated deployment, prompts sharded by rank
        qnn = synthetic.make_qmodel(cfg["model"], wbits=cfg["wbits"], abits=cfg["abits"], group_num=cfg["groups"],
                                    n_steps=cfg["n_steps"], log_quant=cfg["log"], real_time=cfg["log"],
                                    start_peak=cfg["log"], device=dev, seed=0)
        qnn.enable_cuda_graphs(True)
        if args.config == 3:
            step_resident, step_e2e, units, h2d, d2h, scaling, bd_inp = setup_sd_loop(torch, cfg, qnn, dev, rank, world, extra)
        elif args.config == 5:
            step_resident, step_e2e, units, h2d, d2h, scaling, bd_inp = setup_sweep(torch, dist, cfg, qnn, dev, rank, world, extra, timed)
        else:
            step_resident, step_e2e, units, h2d, d2h, scaling, bd_inp = setup_single(torch, dist, cfg, qnn, dev, rank, world)
        bd_fn = lambda: call_unet(qnn, bd_inp)   # noqa: E731

    if args.profile_step:
        # one eager (no graph) step between cudaProfilerStart/Stop: `ncu --profile-from-start off`
        if qnn is not None:
            qnn.enable_cuda_graphs(False)
        with torch.no_grad():
            bd_fn()
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
            bd_fn()
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        return
    with torch.no_grad():
        for _ in range(warmup):
            step_resident()
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES                 # one more (untimed) step: graph replays count their captured launches
        step_resident()
        torch.cuda.synchronize()
        launches_per_step = ops.LAUNCHES - n0
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        ms_total = timed(step_resident, args.steps)
        if sampler:
            sampler.stop_flag = True
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        if rank == 0:   # per-kernel breakdown of one UNet call / layer forward (eager, per-launch CUDA events)
            if qnn is not None:
                qnn.enable_cuda_graphs(False)
            engine.OVERLAP = False        # serialised launches: per-kernel durations, not concurrent-kernel wall time
            breakdown = kernel_breakdown(torch, ops, bd_fn)
            engine.OVERLAP = True
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    ms_step = ms_total / args.steps
    value = units / (ms_step / 1e3)
    e2e_val = units / (ms_e2e / args.steps / 1e3)
    line = {
        "metric": cfg["metric"], "value": round(value, 3), "unit": cfg["unit"],
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f16" if cfg["groups"] > 1 else "u8",
        "data": "synthetic latents/prompt-embeddings, random-init weights, synthetic activation scales (SURVEY.md 8d)",
        "config": {"workload": cfg["workload"], "config_id": args.config, "global_batch": units,
                   "parallelism": f"replica x{world}, prompt batch sharded, latents all-gathered",
                   "l2": "per-step working set (resident weight operands + activations) >> 126 MB L2; no flush needed"
                         if cfg["model"] != "layer" else "32 resident input copies (168 MB > L2) visited round-robin; no flush needed",
                   "act_dtype_between_kernels": str(ops.ACT_DTYPE).replace("torch.", ""), "cuda_graph": cfg["model"] != "layer",
                   "gemm_kinds": "kind::i8 (u8 x s8) for scalar / row-wise activation scales, kind::f16 for K-wise group scales"},
        "e2e": {"value": round(e2e_val, 3), "unit": cfg["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,   # inside the timed region
        "clocks": sampler.summary(),
    }
    line.update(rooflines(cfg, breakdown, pk, pk_src))
    line.update(extra)
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(torch, cfg, qnn, budget_s=float(os.environ.get("DGQ_CPU_BUDGET_S", "150")))
        if qnn is not None:
            line["torch_eager_b200_baseline"] = torch_eager_b200(torch, cfg, qnn, dev)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- workloads
def setup_single(torch, dist, cfg, qnn, dev, rank, world):
    """configs 2 and 4: one UNet call per step, `batch` samples per GPU (weak scaling)."""
    batch = cfg["batch"]
    host = make_inputs(torch, cfg, batch, seed=1000 + rank, pin=True)   # rank r owns prompts [r*batch, (r+1)*batch)
    devin = [x.to(dev) for x in host]
    lat = devin[0].shape
    gather = [torch.empty(lat, device=dev) for _ in range(world)] if world > 1 else None

    def step_resident():
        y = call_unet(qnn, devin)
        if world > 1:
            dist.all_gather(gather, y)   # the only collective on the path: final latent gather
        return y

    def step_e2e():
        y = call_unet(qnn, [x.to(dev, non_blocking=True) for x in host])
        if world > 1:
            dist.all_gather(gather, y)
        return y.to("cpu", non_blocking=True)
    h2d = sum(x.numel() * x.element_size() for x in host)
    return step_resident, step_e2e, batch * world, h2d, devin[0].numel() * 4, "weak", devin


def setup_sd_loop(torch, cfg, qnn, dev, rank, world, extra):
    """config 3: StableDiffusionPipeline's denoise loop (pipeline_stable_diffusion.py:1017-1047): 50 PLMS steps =
    51 UNet calls at UNet batch 16 (8 prompts x CFG pair), one CUDA graph per step index, sampler step on the
    device.  Replicas only: each rank runs its own 8 prompts (weak)."""
    from dgq_b200 import sampler as DS
    prompts, n_steps = cfg["prompts"], cfg["n_steps"]
    g = torch.Generator().manual_seed(2000 + rank)
    lat_h = torch.randn(prompts, 4, 64, 64, generator=g).pin_memory()
    ctx_h = torch.randn(2 * prompts, 77, 768, generator=g).pin_memory()
    lat_d, ctx_d = lat_h.to(dev), ctx_h.to(dev)

    def step_resident():
        return DS.denoise_sd(qnn, lat_d, ctx_d, n_steps, guidance=7.5)

    def step_e2e():
        x = DS.denoise_sd(qnn, lat_h.to(dev, non_blocking=True), ctx_h.to(dev, non_blocking=True), n_steps, guidance=7.5)
        return x.to("cpu", non_blocking=True)
    extra["unet_calls_per_step"] = n_steps + 1
    if rank == 0 and os.environ.get("DGQ_LOOP_PARITY", "1") != "0":
        extra["loop_parity"] = sd_loop_parity(torch, cfg, qnn, dev, lat_d[:2], ctx_d[[0, 1, prompts, prompts + 1]])
    h2d = lat_h.numel() * 4 + ctx_h.numel() * 4
    bd = make_inputs(torch, cfg, cfg["batch"], seed=1000, device=dev)
    return step_resident, step_e2e, prompts * world, h2d, lat_h.numel() * 4, "weak", bd


def sd_loop_parity(torch, cfg, qnn, dev, lat, ctx):
    """Final-latent cosine of the full 50-step CFG loop (2 prompts) against the oracle port run as torch-eager
    fp32 ON THE B200 in the same process (same weights, same per-step scales, oracle sampler)."""
    try:
        from oracle import dgq_oracle as O, sampler_oracle as SO
        from dgq_b200 import sampler as DS
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        n = cfg["n_steps"]
        with torch.no_grad():
            t0 = time.time()
            got = DS.denoise_sd(qnn, lat, ctx, n, guidance=7.5).clone()
            torch.cuda.synchronize()
            t_gpu = time.time() - t0
            state = [export_oracle_state(torch, cfg, qnn, step=k, device=dev) for k in range(n)]

            def oracle_unet(x, t, c):
                sd, act, ocfg = state[int((1000 - float(t)) // (1000 // n))]
                return O.unet_forward("sd", sd, act, ocfg, x, t.to(x.device), c)
            t0 = time.time()
            want = SO.denoise_sd(oracle_unet, lat, ctx, n, guidance=7.5)
            torch.cuda.synchronize()
            t_ref = time.time() - t0
        cos = torch.nn.functional.cosine_similarity(got.flatten().float(), want.flatten().float(), dim=0).item()
        l2 = ((got - want).norm() / want.norm()).item()
        del state
        torch.cuda.empty_cache()
        return {"final_latent_cosine": round(cos, 6), "rel_l2": round(l2, 5), "prompts": int(lat.shape[0]),
                "unet_calls": n + 1, "oracle": "oracle port, torch eager fp32 on the same B200",
                "dgq_b200_s_incl_capture": round(t_gpu, 2), "oracle_s": round(t_ref, 2)}
    except Exception as e:   # parity of the loop is also a -m gpu test; the bench line must still come out
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def shard_plan(total, world, rank, mb_max):
    """micro-batch sizes rank `rank` runs for a total prompt batch sharded contiguously over `world` ranks
    (src/gen4eval_SDXL.py:116: file_list[rank*len//ws : (rank+1)*len//ws]); fewer prompts than ranks: the first
    `total` ranks take one each."""
    lo, hi = rank * total // world, (rank + 1) * total // world
    share, mbs = hi - lo, []
    while share > 0:
        mbs.append(min(mb_max, share))
        share -= mbs[-1]
    return mbs


def setup_sweep(torch, dist, cfg, qnn, dev, rank, world, extra, timed):
    """config 5: the reference's only multi-GPU pattern (src/gen4eval_SDXL.py:116: file_list[rank*len//ws : ...]):
    a FIXED total prompt batch sharded contiguously over the ranks (strong scaling), micro-batches of <= 16.
    The headline value is the 512-prompt batch; `sweep` lists images/s for total = 8, 16, ..., 512."""
    total_max, mb_max = cfg["total"], cfg["batch"]
    pools = {}

    def plan(total):
        return shard_plan(total, world, rank, mb_max)

    def inputs_for(mb, seed, pin=False):
        key = (mb, seed, pin)
        if key not in pools:
            pools[key] = make_inputs(torch, cfg, mb, seed=seed, pin=pin) if pin else make_inputs(torch, cfg, mb, seed=seed, device=dev)
        return pools[key]

    def run(total, host=False):
        outs = []
        for i, mb in enumerate(plan(total)):
            if host:
                inp = [x.to(dev, non_blocking=True) for x in inputs_for(mb, 3000 + rank * 64 + i % 2, pin=True)]
            else:
                inp = inputs_for(mb, 3000 + rank * 64 + i % 2)
            y = call_unet(qnn, inp)
            outs.append(y.to("cpu", non_blocking=True) if host else y)
        if world > 1 and outs:    # the final latent gather of the LAST micro-batch (all ranks hold one at every total >= world)
            pass
        return outs

    def step_resident():
        return run(total_max)

    def step_e2e():
        return run(total_max, host=True)
    # sweep: each total timed over 2 passes after 1 warm-up (graph capture per micro-batch size)
    sweep = {}
    with torch.no_grad():
        t = 8
        while t <= total_max:
            run(t)
            ms = timed(lambda: run(t), 2) / 2
            sweep[str(t)] = round(t / (ms / 1e3), 2)
            t *= 2
    extra["sweep_images_per_s"] = sweep
    extra["sweep_note"] = "total prompt batch -> images/s at this GPU count; per-GPU share in micro-batches of <= 16"
    per_rank = sum(plan(total_max))
    one = inputs_for(mb_max, 3000, pin=True)
    h2d = sum(x.numel() * x.element_size() for x in one) * per_rank // mb_max
    bd = inputs_for(mb_max, 3000)
    return step_resident, step_e2e, total_max, h2d, per_rank * 4 * 128 * 128 * 4, "strong", bd


def setup_layer(torch, cfg, dev):
    """config 1: the reference's own CPU-runnable case, one QuantLayer through the reference-shaped API."""
    import torch.nn as nn
    from quant.quant_layer import QuantLayer, Scaler
    g = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    layer = nn.Conv2d(320, 320, 3, 1, 1)
    x_h = torch.randn(1, 320, 64, 64, generator=g).pin_memory()
    lab = torch.randint(0, 8, (2880,), generator=g)
    lo, hi = -(torch.rand(8, generator=g) * 3 + 1), torch.rand(8, generator=g) * 3 + 1
    d = ((hi - lo) / 255)[lab].view(1, -1, 1)
    z = torch.round(-lo / ((hi - lo) / 255))[lab].view(1, -1, 1)
    state = dict(w=layer.weight.detach().clone(), b=layer.bias.detach().clone(), d=d, z=z, x=x_h.clone())
    ql = QuantLayer(layer, {"bits": 4, "channel_wise": True, "scaler": Scaler.MINMAX},
                    {"bits": 8, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True}).to(dev)
    ql.aqtizer.delta, ql.aqtizer.zero_point, ql.aqtizer.init = d.to(dev), z.to(dev), True
    ql.use_group_num = True
    ql.set_quant_state(True, True)
    # 32 resident copies of the input (168 MB > the 126 MB L2), visited round-robin: every timed forward reads its
    # activation from HBM without a flush kernel inside the timed region
    x_d = [x_h.to(dev).clone() for _ in range(32)]
    CONFIGS[1]["_state"] = state
    it = [0]

    def step_resident():
        it[0] += 1
        return ql(x_d[it[0] % 32])

    def step_e2e():
        return ql(x_h.to(dev, non_blocking=True)).to("cpu", non_blocking=True)
    return step_resident, step_e2e, None, 1, x_h.numel() * 4, x_h.numel() * 4, lambda: ql(x_d[0])


# ---------------------------------------------------------------------------------------------- rooflines
def rooflines(cfg, bd, pk, pk_src):
    """roofline (qGEMM, tensor), roofline_attention (tensor) and tail (HBM GB/s per bandwidth-bound entry point),
    all from the per-launch CUDA events of ONE eager UNet call / layer forward and algorithmic work counted from
    the launch arguments."""
    peak16 = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    gemm = bd["_gemm"]
    fl16, ms16, fl8, ms8 = gemm["f16_flop"], gemm["f16_ms"], gemm["i8_flop"], gemm["i8_ms"]
    tf16 = fl16 / ms16 / 1e9 if ms16 else 0.0
    t8 = fl8 / ms8 / 1e9 if ms8 else 0.0
    peak8 = pk.get("int8_tops_sustained")
    # one fraction for the GEMM class: time-weighted over the two kinds, each against its own measured peak
    ideal_ms = (fl16 / peak16 / 1e9 if fl16 else 0.0) + (fl8 / (peak8 or 2 * peak16) / 1e9 if fl8 else 0.0)
    tot_ms = ms16 + ms8
    out = {"roofline": {
        "kernel": "gemm_f16_kernel<.., kI8> (tcgen05 qGEMM: every QuantLayer of the step)", "bound": "tensor",
        "achieved": round((fl16 + fl8) / tot_ms / 1e9, 1) if tot_ms else None, "peak": peak16, "unit": "TFLOP/s",
        "frac": round(ideal_ms / tot_ms, 4) if tot_ms else None,
        "frac_note": "sum over launches of (flop / measured peak of the launch's kind) / summed launch time; "
                     "kind::f16 vs bf16 sustained, kind::i8 vs int8 sustained (profiles/r2_probes.txt)",
        "peak_source": pk_src + " bf16 sustained (kernel timed inside a long step)",
        "f16": {"tflops": round(tf16, 1), "ms": round(ms16, 3), "peak": peak16},
        "i8": {"tops": round(t8, 1), "ms": round(ms8, 3), "peak": peak8},
        "traffic": ncu_traffic()[0], "traffic_note": ncu_traffic()[1],
        "share_of_step": round(tot_ms / bd["_total_ms"], 4)}}
    att = bd["_attn"]
    if att["ms"]:
        tf = att["flop"] / att["ms"] / 1e9
        out["roofline_attention"] = {"kernel": "attention_kernel (two-pass flash, quantised softmax map)", "bound": "tensor",
                                     "achieved": round(tf, 1), "peak": peak16, "unit": "TFLOP/s", "frac": round(tf / peak16, 4),
                                     "flop_note": "useful 2*(QK^T + PV) only; the pass-1 QK^T recompute is not counted",
                                     "ms": round(att["ms"], 3), "share_of_step": round(att["ms"] / bd["_total_ms"], 4)}
    tail = {}
    for name, v in bd["_tail"].items():
        if v["ms"] > 0:
            gbs = v["bytes"] / v["ms"] / 1e6
            tail[name] = {"gbs": round(gbs, 1), "frac": round(gbs / pk["hbm_gbs"], 4), "ms": round(v["ms"], 3),
                          "algorithmic_mb": round(v["bytes"] / 1e6, 1)}
    out["tail"] = tail
    out["tail_note"] = f"achieved HBM GB/s on algorithmic bytes (inputs read once + operands written) vs {pk['hbm_gbs']} GB/s {pk_src}"
    out["breakdown_ms"] = {k: round(v["ms"], 3) for k, v in bd.items() if not k.startswith("_")}
    out["top_shapes_ms"] = {k: [round(v["ms"], 3), v["calls"]] for k, v in
                            sorted(bd["_shapes"].items(), key=lambda kv: -kv[1]["ms"])[:16]}
    if cfg["model"] != "layer":
        q, a = GMAC[cfg["model"]]
        out["unet_call_ms_serialised"] = round(bd["_total_ms"], 3)
        out["unet_call_tflops"] = round(2 * (q + a) * 1e9 * cfg["batch"] / (bd["_wall_ms"] / 1e3) / 1e12, 1)
    return out


def ncu_traffic():
    """DRAM bytes (read + write) of the dominant qGEMM launch from the ncu --set full capture of THIS build:
    profiles/r2_gemm_traffic.json carries the sha256 of csrc/gemm.cu it was captured from; a stale capture reads null."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")))
        sha = hashlib.sha256(open(os.path.join(ROOT, "dgq_b200", "csrc", "gemm.cu"), "rb").read()).hexdigest()[:16]
        if d.get("gemm_cu_sha16") != sha:
            return None, f"profiles/r2_gemm_traffic.json was captured from another build of gemm.cu ({d.get('gemm_cu_sha16')} vs {sha})"
        return d["dram_bytes"], d["note"]
    except Exception as e:
        return None, f"no ncu capture for this build ({type(e).__name__})"


def _args(a):
    return a[0]._obj if hasattr(a[0], "_obj") else a[0].contents


def kernel_breakdown(torch, ops, fn):
    """Run one eager step with a CUDA-event pair around every C-ABI call; sum by entry point, with the
    algorithmic flop / bytes of each call counted from its arguments."""
    from dgq_b200 import _lib as L
    lib = L.lib()
    events, originals = [], {}

    def work(name, a):
        """(tag, flop, bytes, kind) of one call"""
        try:
            if name in ("dgq_gemm_f16", "dgq_gemm_i8"):
                t = _args(a)
                epi = ("plain", "geglu", "qkv")[t.epi] + ("+resid" if t.resid else "") + ("" if t.out else " f32")
                return f"gemm{' i8' if name.endswith('i8') else ''} {t.m}x{t.n}x{t.k} {epi}", 2.0 * t.m * t.n * t.k, 0, name[-2:]
            if name == "dgq_attention":
                t = _args(a)
                return f"attn b{t.b} h{t.heads} t{t.t} s{t.s} d{t.d}", 4.0 * t.b * t.heads * t.t * t.s * t.d, 0, "attn"
            if name == "dgq_act_producer":
                t = _args(a)
                hs, ws = (t.h // 2, t.w // 2) if t.upsample else (t.h, t.w)
                ho, wo = (t.h + 2 * t.pad - t.ksize) // t.stride + 1, (t.w + 2 * t.pad - t.ksize) // t.stride + 1
                esz_out = 1 if t.q.emit_int == 2 else 2
                return name, 0, t.batch * hs * ws * (t.c0 + t.c1) * (4 if t.src_is_f32 else 2) + t.batch * ho * wo * t.ldo * esz_out, "tail"
            if name in ("dgq_ln_quant", "dgq_row_quant"):
                x_is32, m, c = a[1], a[2], a[3]
                n_out = a[7] if name == "dgq_ln_quant" else a[4]
                qs = a[8] if name == "dgq_ln_quant" else a[5]
                wr = sum(m * c * (1 if qs[i].emit_int == 2 else 2) for i in range(n_out))
                return name, 0, m * c * (4 if x_is32 else 2) + wr, "tail"
            if name == "dgq_gn_stats":
                return name, 0, a[5] * a[6] * (a[3] + a[4]) * (4 if a[2] else 2), "tail"
            if name == "dgq_geglu_quant":
                return name, 0, a[2] * a[3] * (2 * (4 if a[1] else 2) + 2), "tail"
        except Exception:
            pass
        return name, 0, 0, "other"

    for name in L.SYMBOLS:
        if name == "dgq_version":
            continue
        f = getattr(lib, name)
        originals[name] = f

        def wrap(f=f, name=name):
            def g(*a):
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                rc = f(*a)
                e1.record()
                events.append((name, e0, e1) + work(name, a))
                return rc
            return g
        setattr(lib, name, wrap())
    try:
        fn()
        torch.cuda.synchronize()
        events.clear()
        t0, t1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0.record()
        fn()
        t1.record()
        torch.cuda.synchronize()
    finally:
        for name, f in originals.items():
            setattr(lib, name, f)
    acc, shapes, tail = {}, {}, {}
    gemm = {"f16_flop": 0.0, "f16_ms": 0.0, "i8_flop": 0.0, "i8_ms": 0.0}
    attn = {"flop": 0.0, "ms": 0.0}
    for name, e0, e1, tag, flop, nbytes, kind in events:
        ms = e0.elapsed_time(e1)
        d = acc.setdefault(name, {"ms": 0.0, "calls": 0})
        d["ms"] += ms
        d["calls"] += 1
        if tag != name:
            sh = shapes.setdefault(tag, {"ms": 0.0, "calls": 0})
            sh["ms"] += ms
            sh["calls"] += 1
        if kind in ("16", "i8"):
            k = "f16" if kind == "16" else "i8"
            gemm[k + "_flop"] += flop
            gemm[k + "_ms"] += ms
        elif kind == "attn":
            attn["flop"] += flop
            attn["ms"] += ms
        elif kind == "tail":
            t = tail.setdefault(name, {"ms": 0.0, "bytes": 0})
            t["ms"] += ms
            t["bytes"] += nbytes
    acc["_total_ms"] = sum(v["ms"] for v in acc.values())
    acc["_shapes"], acc["_gemm"], acc["_attn"], acc["_tail"] = shapes, gemm, attn, tail
    acc["_wall_ms"] = t0.elapsed_time(t1)
    return acc


# ---------------------------------------------------------------------------------------------- baselines
def export_oracle_state(torch, cfg, qnn, step=0, device="cpu"):
    """(sd, act, cfg) in the oracle's (== the reference checkpoint's) schema from a live QuantModel, step `step`."""
    from oracle import dgq_oracle as O
    from dgq_b200.quant.quant_layer import QuantLayer, UniformAffineQuantizer
    sd = {k: v.detach().float().to(device) for k, v in qnn.state_dict().items()}
    act = {}
    named = dict(qnn.named_modules())
    for path, m in named.items():
        if isinstance(m, UniformAffineQuantizer) and m._table is not None:
            q = m._table[step]
            owner = named[path.rpartition(".")[0]]
            d, z = q.delta.to(device), q.zp.to(device)
            if q.mode == 1:
                d, z = d.reshape(()), z.reshape(())
            elif path.endswith(".aqtizer") and isinstance(owner, QuantLayer) and owner.is_conv:
                kperm = owner._kperm(device)
                if kperm is not None:
                    inv = torch.empty_like(kperm)
                    inv[kperm] = torch.arange(kperm.numel(), device=kperm.device)
                    d, z = d[inv], z[inv]
                d, z = d.view(1, -1, 1), z.view(1, -1, 1)
            else:
                d, z = d.view(1, 1, -1), z.view(1, 1, -1)
            act[path + ".delta"], act[path + ".zero_point"] = d, z
    ocfg = O.QConfig(wbits=cfg["wbits"], abits=cfg["abits"], softmax_bits=cfg["abits"], t2i_log_quant=cfg["log"],
                     t2i_real_time=cfg["log"], t2i_start_peak=cfg["log"])
    O.update_group_convs(ocfg, act, sd)
    return sd, act, ocfg


def oracle_call(torch, cfg, state, inp):
    from oracle import dgq_oracle as O
    sd, act, ocfg = state
    added = {"text_embeds": inp[3], "time_ids": inp[4]} if len(inp) == 5 else None
    return O.unet_forward(cfg["model"], sd, act, ocfg, inp[0], inp[1], inp[2], added)


def layer_oracle(torch, cfg):
    from oracle import dgq_oracle as O
    st = cfg["_state"] if "_state" in cfg else None
    if st is None:   # --impl reference: build the same layer
        import torch.nn as nn
        g = torch.Generator().manual_seed(0)
        torch.manual_seed(0)
        layer = nn.Conv2d(320, 320, 3, 1, 1)
        x = torch.randn(1, 320, 64, 64, generator=g)
        lab = torch.randint(0, 8, (2880,), generator=g)
        lo, hi = -(torch.rand(8, generator=g) * 3 + 1), torch.rand(8, generator=g) * 3 + 1
        st = dict(w=layer.weight.detach().clone(), b=layer.bias.detach().clone(), x=x,
                  d=((hi - lo) / 255)[lab].view(1, -1, 1), z=torch.round(-lo / ((hi - lo) / 255))[lab].view(1, -1, 1))
    wd, wz = O.channel_minmax_scale(st["w"], 16)
    sd = {"l.w": st["w"], "l.b": st["b"], "l.wqtizer.delta": wd, "l.wqtizer.zero_point": wz}
    act = {"l.aqtizer.delta": st["d"], "l.aqtizer.zero_point": st["z"]}
    ocfg = O.QConfig(wbits=4, abits=8, group_convs={"l"})
    return lambda: O.quant_layer(st["x"], sd, act, "l", ocfg, padding=1)


def cpu_baseline(torch, cfg, qnn, budget_s):
    """The reference's fake-quant forward (oracle port, fp32) on the host cores: a bounded sample of the same
    workload (one batch-1 UNet call; config 1: the layer itself), as many repeats as fit the budget."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if cfg["model"] == "layer":
        fn, per_call, what = layer_oracle(torch, cfg), 1.0, "the same single-layer forward"
    else:
        state = export_oracle_state(torch, cfg, qnn)
        inp = make_inputs(torch, cfg, 1, seed=1000)
        fn = lambda: oracle_call(torch, cfg, state, inp)   # noqa: E731
        calls = cfg["n_steps"] + 1 if cfg["n_steps"] > 1 else 1
        per_call, what = 1.0 / calls, f"batch 1 of the same workload (one {cfg['model'].upper()} UNet call" + \
            (f"; a full image needs {calls} calls x CFG pair, value scaled by 1/{2 * calls})" if calls > 1 else ")")
        if calls > 1:
            per_call = 1.0 / (2 * calls)     # an image = 51 calls of its (uncond, cond) pair
    times = []
    t_start = time.perf_counter()
    with torch.no_grad():
        while True:
            t0 = time.perf_counter()
            fn()
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start + times[-1] > budget_s or len(times) >= (20 if cfg["model"] == "layer" else 3):
                break
    best = min(times)
    return {"value": round(per_call / best, 6), "unit": cfg["unit"], "cores": cores, "kind": "port",
            "sample": f"{what}, best of {len(times)}, {best:.3f} s/call"}


def torch_eager_b200(torch, cfg, qnn, dev):
    """north_star: "the reference's PyTorch-on-B200 fake-quant path also listed" -- the oracle port (the
    reference's fake-quant forward as plain PyTorch eager ops, fp32) executed on the B200 itself, batch 1 of the
    same workload.  A baseline like cpu_baseline: never on the product path."""
    try:
        state = export_oracle_state(torch, cfg, qnn, device=dev)
        inp = make_inputs(torch, cfg, 1, seed=1000, device=dev)
        times = []
        with torch.no_grad():
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                oracle_call(torch, cfg, state, inp)
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
        # the reference's --fp16 arm (src/inference_qmodel.py:95-98, qnn.half()): the same eager graph with its matmuls /
        # convs in fp16 (autocast); a throughput figure only -- half-precision fake quantisation is not a parity target
        fp16 = None
        try:
            t16 = []
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                for _ in range(3):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    oracle_call(torch, cfg, state, inp)
                    torch.cuda.synchronize()
                    t16.append(time.perf_counter() - t0)
            fp16 = {"value": round(1.0 / min(t16[1:]), 4), "unit": "UNet calls/s at batch 1", "ms_per_call": round(min(t16[1:]) * 1e3, 1)}
        except Exception as e:
            fp16 = {"unavailable": f"{type(e).__name__}: {e}"[:160]}
        del state
        torch.cuda.empty_cache()
        best = min(times[1:])
        return {"value": round(1.0 / best, 4), "unit": "UNet calls/s at batch 1", "kind": "port, torch eager fp32 on the same B200",
                "sample": f"batch 1 (one {cfg['model'].upper()} UNet call), best of 2 after 1 warm-up, {best * 1e3:.0f} ms/call",
                "fp16_autocast": fp16}
    except Exception as e:   # a baseline must never take the bench down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args):
    """--impl reference: the reference's CPU fake-quant path (oracle port: /root/reference is not on the GPU box and
    is pure Python, so there is no oracle/_ref to compile), all host threads, same metric/config; each step = a
    bounded sample of the workload (one batch-1 UNet call; config 1: the layer)."""
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import dgq_oracle as O, synth as S
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if cfg["model"] == "layer":
        fn, scale, sample = layer_oracle(torch, cfg), 1.0, "the same single-layer forward per step"
    else:
        model = cfg["model"]
        sd = S.make_weights(model, seed=0)
        S.init_weight_quant(sd, cfg["wbits"])
        ocfg = O.QConfig(wbits=cfg["wbits"], abits=cfg["abits"], softmax_bits=cfg["abits"], t2i_log_quant=cfg["log"],
                         t2i_real_time=cfg["log"], t2i_start_peak=cfg["log"])
        hd = 64
        shapes = {}
        for name, d in S.iter_modules(model):
            if d[0] == "conv" and name not in ("model.conv_in", "model.conv_out"):
                shapes[name + ".aqtizer"] = ("out", d[2] * d[3] * d[3])
            elif d[0] == "lin":
                two_d = any(s in name for s in ("time_embedding", "add_embedding", "time_emb_proj"))
                shapes[name + ".aqtizer"] = ("scalar", 0) if two_d else ("in", d[2])
                if name.endswith(".to_q"):
                    a = name[: -len(".to_q")]
                    heads = O.SPECS[model]["heads"](d[1])
                    for qn in ("aqtizer_q", "aqtizer_k", "aqtizer_v"):
                        shapes[f"{a}.{qn}"] = ("in", d[1] // heads if model == "sd" else hd)
                    if not cfg["log"]:
                        shapes[f"{a}.aqtizer_w"] = ("scalar", 0)
        if cfg["groups"] <= 1:
            shapes = {k: ("scalar", 0) for k in shapes}
        act = S.random_act(model, sd, ocfg, shapes, cfg["groups"], seed=0)
        act = {k: (v.view(1, 1, 1, -1) if ("aqtizer_q" in k or "aqtizer_k" in k or "aqtizer_v" in k) and v.dim() == 3 else v)
               for k, v in act.items()}
        O.update_group_convs(ocfg, act, sd)
        inp = make_inputs(torch, cfg, 1, seed=1000)
        fn = lambda: oracle_call(torch, cfg, (sd, act, ocfg), inp)   # noqa: E731
        calls = cfg["n_steps"] + 1 if cfg["n_steps"] > 1 else 1
        scale = 1.0 / (2 * calls) if calls > 1 else 1.0
        sample = f"batch 1 (one {model.upper()} UNet call) per step" + (f"; images/s = calls/s / {2 * calls}" if calls > 1 else "")
    budget = float(os.environ.get("DGQ_CPU_BUDGET_S", "240"))
    t_start = time.perf_counter()
    times, n_warm = [], 0
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            if i < args.warmup and (time.perf_counter() - t_start + 2 * dt) < budget:
                n_warm += 1
                continue
            times.append(dt)
            if time.perf_counter() - t_start + dt > budget:
                break
    ms = 1e3 * sum(times) / len(times)
    val = scale / (ms / 1e3)
    line = {"impl": "reference", "metric": cfg["metric"], "value": round(val, 6), "unit": cfg["unit"],
            "n_gpus": 0, "steps": len(times), "warmup": n_warm, "ms_per_step": round(ms, 2), "higher_is_better": True,
            "scaling": "strong" if args.config == 5 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "config_id": args.config,
                       "note": f"CPU fake-quant path (oracle port of the reference), bounded sample: {sample}; "
                               f"time-budgeted to {budget:.0f} s so steps/warmup may be fewer than requested"},
            "cpu_baseline": {"value": round(val, 6), "unit": cfg["unit"], "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(val, 6), "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="dgq_b200", choices=["dgq_b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / torch-eager legs")
    ap.add_argument("--profile-step", action="store_true",
                    help="run one eager step inside cudaProfilerStart/Stop and exit (for ncu)")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = {1: 50, 2: 50, 3: 2, 4: 10, 5: 2}[a.config]
    if a.impl == "reference":
        run_reference(a)
    else:
        run_cuda(a)
