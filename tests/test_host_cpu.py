"""CPU-side checks: the C-ABI library loads and exports every symbol include/dgq_b200.h declares
(no compute calls without a GPU), the drop-in import paths and class identities hold, the module
tree produces the reference's state-dict keys, the loader's time-aware tables / sticky
use_group_num logic, and the 'no CPU fallback' contract."""
import os
import re
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dgq_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dgq_b200.h")).read()
    declared = set(re.findall(r"^int (dgq_\w+)\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dgq_version() >= 100


def test_ctypes_structs_match_header_layout():
    """field order/count of the ctypes mirrors vs the C structs (a silent mismatch corrupts launches)."""
    from dgq_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dgq_b200.h")).read()

    def c_fields(struct_name):
        end = hdr.index("} " + struct_name + ";")
        body = hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            parts = re.sub(r"\[\d+\]", "", decl).replace("*", " ").split(",")
            first = parts[0].split()
            names.append(first[-1])
            names += [p.strip().split()[-1] for p in parts[1:]]
        return names

    for cname, ct in (("dgq_quant_t", _lib.QuantT), ("dgq_producer_t", _lib.ProducerT),
                      ("dgq_gemm_t", _lib.GemmT), ("dgq_attn_t", _lib.AttnT),
                      ("dgq_sampler_step_t", _lib.SamplerStepT)):
        assert c_fields(cname) == [f[0] for f in ct._fields_], cname


def test_drop_in_import_paths_and_identity():
    import quant.quant_model, quant.quant_layer, quant.quant_block, quant.quant_layer_text  # noqa
    import quant.load_qmodel_util, quant.calibration  # noqa
    import dgq_b200.quant.quant_model as impl
    from quant.quant_model import QuantModel
    assert QuantModel is impl.QuantModel
    from quant.quant_layer import Scaler, QMODE, QuantLayer, UniformAffineQuantizer, StraightThrough  # noqa
    from quant.quant_block import BaseQuantBlock, QuantBasicTransformerBlock, QuantResnetBlock2D  # noqa
    from quant.quant_layer_text import T2ILogQuantizer  # noqa
    from quant.load_qmodel_util import get_qmodel  # noqa
    from quant.calibration import load_cali_model  # noqa
    import diffusers_rewrite
    for n in ("UNet2DConditionModel", "Attention", "ResnetBlock2D", "BasicTransformerBlock", "Timesteps",
              "TimestepEmbedding"):
        assert hasattr(diffusers_rewrite, n)
    assert callable(Scaler.MINMAX)


WQ = {"bits": 4, "channel_wise": True}
AQ = {"bits": 8, "channel_wise": False, "leaf_param": True}
SM = {"softmax_a_bit": 8, "t2i_log_quant": True, "t2i_real_time": True, "t2i_start_peak": True, "log_max_1": False}


def _meta_qmodel(model_type):
    from quant.quant_model import QuantModel
    from quant.quant_layer import Scaler
    from dgq_b200.unet import sd, sdxl
    graph = sdxl if model_type == "sdxl" else sd
    with torch.device("meta"):
        unet = graph.UNet2DConditionModel()
        return QuantModel(unet, dict(WQ, scaler=Scaler.MINMAX), dict(AQ, scaler=Scaler.MINMAX), SM)


@pytest.mark.parametrize("model_type,n_layers,n_keys", [("sd", 282, 686), ("sdxl", 794, 1680)])
def test_module_tree_matches_reference_schema(model_type, n_layers, n_keys):
    from quant.quant_layer import QuantLayer
    from quant.quant_block import QuantBasicTransformerBlock, QuantResnetBlock2D
    from oracle import synth
    q = _meta_qmodel(model_type)
    assert sum(isinstance(m, QuantLayer) for m in q.modules()) == n_layers        # SURVEY.md 3.3 [probe]
    keys = set(q.state_dict().keys())
    want = set()
    for n, d in synth.iter_modules(model_type):
        if d[0] in ("gn", "ln"):
            want |= {n + ".weight", n + ".bias"}
        else:
            want.add(n + ".w")
            if d[0] == "conv" or d[3]:
                want.add(n + ".b")
    assert keys == want and len(keys) == n_keys
    blocks = [m for m in q.modules() if isinstance(m, QuantBasicTransformerBlock)]
    assert all(hasattr(b.attn2, "start_peak") and not hasattr(b.attn1, "start_peak") for b in blocks)
    assert q.config.in_channels == 4 and q.config.sample_size == (128 if model_type == "sdxl" else 64)
    # set_quant_state / disable_out_quantization semantics (reference quant_model.py:105-124)
    q.set_quant_state(True, True)
    q.disable_out_quantization()
    assert q.model.conv_in.use_wq is False and q.model.conv_in.disable_aq is True
    assert all(b.attn1.use_aq for b in blocks)
    res = next(m for m in q.modules() if isinstance(m, QuantResnetBlock2D))
    res.conv1.ignore_recon = True
    q.set_quant_state(True, True)
    assert res.conv1.use_wq is False


def test_step_index_and_sticky_group_flag():
    from quant.quant_layer import QuantLayer
    from dgq_b200.quant.quant_model import QuantModel
    q = _meta_qmodel("sd")
    q._num_inference_steps = 50
    assert q.step_index(torch.tensor(981)) == 0 and q.step_index(torch.tensor([961])) == 1
    assert q.step_index(torch.tensor(1)) == 49
    q._num_inference_steps = 4
    assert [q.step_index(torch.tensor(t)) for t in (999, 749, 499, 249)] == [0, 1, 2, 3]


def test_qparam_axis_mapping():
    """(1,1,X)/(1,X,1) checkpoint shapes -> K-wise / row-wise, swapped for unfolded conv inputs."""
    from dgq_b200 import ops
    d3, z3 = torch.rand(1, 1, 12) + 0.1, torch.zeros(1, 1, 12)
    assert ops.qparam_from_ckpt(d3, z3, 255.0, "cpu").mode == ops.Q_KWISE
    assert ops.qparam_from_ckpt(d3, z3, 255.0, "cpu", conv=True).mode == ops.Q_ROWWISE
    d2, z2 = d3.view(1, 12, 1), z3.view(1, 12, 1)
    assert ops.qparam_from_ckpt(d2, z2, 255.0, "cpu").mode == ops.Q_ROWWISE
    kperm = torch.arange(12).flip(0)
    q = ops.qparam_from_ckpt(d2, z2, 255.0, "cpu", conv=True, kperm=kperm)
    assert q.mode == ops.Q_KWISE and torch.equal(q.delta, d2.reshape(-1)[kperm])
    s = ops.qparam_from_ckpt(torch.tensor(0.1), torch.tensor(3.0), 255.0, "cpu")
    assert s.mode == ops.Q_SCALAR and s.exact
    assert not ops.qparam_from_ckpt(torch.tensor(0.1), torch.tensor(4000.0), 255.0, "cpu").exact
    with pytest.raises(ValueError):
        ops.qparam_from_ckpt(torch.rand(2, 3, 4), torch.rand(2, 3, 4), 255.0, "cpu")


def test_no_cpu_fallback():
    from quant.quant_layer import QuantLayer, UniformAffineQuantizer, Scaler
    from quant.quant_layer_text import T2ILogQuantizer
    ql = QuantLayer(torch.nn.Linear(8, 8), dict(WQ, scaler=Scaler.MINMAX), dict(AQ, scaler=Scaler.MINMAX))
    with pytest.raises(RuntimeError, match="CUDA"):
        ql(torch.zeros(2, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        T2ILogQuantizer(real_time=True)(torch.rand(4))
    from dgq_b200.unet import sd
    with pytest.raises(NotImplementedError):
        sd.ResnetBlock2D(32, 32)(torch.zeros(1, 32, 4, 4), torch.zeros(1, 1280))
    for name in ("MSE", "KL", "HIST", "OMSE"):                      # calibration stays in the reference
        with pytest.raises(NotImplementedError):
            getattr(Scaler, name)(torch.zeros(3))


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    # every rank owns a contiguous slice of the prompt batch, seeded by global position
    host = bench.make_inputs(torch, bench.CONFIGS[4], 2, seed=1000 + rank)
    y = host[0][:, :, :2, :2].contiguous()            # stand-in for this rank's predicted latents
    gather = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(gather, y)
    ms = torch.tensor([float(rank + 1)])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save({"g": gather, "ms": ms}, out)
    dist.destroy_process_group()


def test_sharded_gather_world2_gloo(tmp_path):
    """N>1 host logic on CPU: contiguous prompt sharding, final latent all-gather, max-over-ranks time."""
    import torch.multiprocessing as mp
    import bench
    out = str(tmp_path / "g.pt")
    mp.spawn(_gloo_worker, args=(2, 29631, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ms"].item() == 2.0
    for rank in range(2):
        want = bench.make_inputs(torch, bench.CONFIGS[4], 2, seed=1000 + rank)[0][:, :, :2, :2]
        assert torch.equal(r["g"][rank], want)
    assert not torch.equal(r["g"][0], r["g"][1])


def test_strong_scaling_shard_plan():
    """config 5 (src/gen4eval_SDXL.py:116): a fixed prompt batch split contiguously over the ranks, micro-batches <= 16;
    every prompt is run exactly once whatever the GPU count."""
    import bench
    for total in (8, 16, 64, 512, 100):
        for world in (1, 2, 4, 8):
            shares = [sum(bench.shard_plan(total, world, r, 16)) for r in range(world)]
            assert sum(shares) == total and max(shares) - min(shares) <= 1
            assert all(mb <= 16 for r in range(world) for mb in bench.shard_plan(total, world, r, 16))
    assert bench.shard_plan(512, 8, 3, 16) == [16, 16, 16, 16]
    assert [bench.shard_plan(4, 8, r, 16) for r in range(8)].count([1]) == 4


def test_compiled_header_rejects_foreign_files(tmp_path):
    """compiled-checkpoint reader (host logic only): magic / format checks"""
    from dgq_b200 import compiled
    p = tmp_path / "x.dgqb"
    p.write_bytes(b"not a checkpoint at all" * 4)
    with pytest.raises(ValueError):
        compiled.read_header(str(p))
    import json, struct
    hj = json.dumps({"format": 99}).encode()
    p.write_bytes(compiled.MAGIC + struct.pack("<Q", len(hj)) + hj)
    with pytest.raises(ValueError):
        compiled.read_header(str(p))


def test_cli_accepts_the_reference_flags(monkeypatch):
    """scripts/inference_qmodel.py takes every flag of the reference's src/inference_qmodel.py:17-44 (argument
    names and types), so a command line written for the reference parses unchanged."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("inference_qmodel_cli", os.path.join(ROOT, "scripts", "inference_qmodel.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    argv = ["prog", "--use_group", "--num_inference_steps", "4", "--cali_ckpt", "x.pth", "--fp16", "--wq", "4", "--use_aq",
            "--aq", "8", "--seed", "42", "--t2i_log_quant", "--t2i_real_time", "--t2i_start_peak", "--time_aware_aqtizer"]
    monkeypatch.setattr("sys.argv", argv)
    opt = cli.parse_args()
    assert (opt.use_group, opt.num_inference_steps, opt.cali_ckpt, opt.fp16, opt.wq, opt.use_aq, opt.aq, opt.seed) == \
           (True, 4, "x.pth", True, 4, True, 8, 42)
    assert opt.t2i_log_quant and opt.t2i_real_time and opt.t2i_start_peak and opt.time_aware_aqtizer
    monkeypatch.setattr("sys.argv", ["prog"])
    d = cli.parse_args()       # the reference's defaults
    assert (d.wq, d.aq, d.seed, d.num_inference_steps, d.use_aq, d.use_group) == (4, 8, 42, -1, False, False)


def test_vae_decoder_schema_matches_the_reference_state_dict():
    """dgq_b200.vae.VaeDecoder holds exactly the `post_quant_conv.*` / `decoder.*` keys and shapes of the reference's
    AutoencoderKL (enumerated by oracle.vae_oracle, which tests/golden/make_golden.py loads into the reference)."""
    from dgq_b200.vae import VaeDecoder
    from oracle import vae_oracle as V
    for name in ("sd", "small"):
        cfg = V.VAE_CONFIGS[name]
        ref = V.make_vae_state(cfg, 0)
        mine = VaeDecoder(cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"]).state_dict()
        assert set(mine) == set(ref)
        assert all(tuple(mine[k].shape) == tuple(ref[k].shape) for k in ref)
