"""The reference-facing module API (quant.quant_layer.QuantLayer, QuantBasicTransformerBlock,
UniformAffineQuantizer, T2ILogQuantizer) executed by the CUDA kernels, against outputs the
REFERENCE produced for the same constructor arguments and inputs (tests/golden/ops.pt).
These read like the reference's own usage: build the torch layer, wrap it, set_quant_state, call."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return ((a.float().cpu() - b).abs().max() / b.abs().max()).item()


def make_layer(layer, c, wbits, grouped):
    from quant.quant_layer import QuantLayer, Scaler
    layer.weight.data.copy_(c["weight"]); layer.bias.data.copy_(c["bias"])
    ql = QuantLayer(layer, {"bits": wbits, "channel_wise": True, "scaler": Scaler.MINMAX},
                    {"bits": 8, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True}).to(DEV)
    ql.aqtizer.delta, ql.aqtizer.zero_point, ql.aqtizer.init = c["delta"].to(DEV), c["zp"].to(DEV), True
    ql.use_group_num = grouped
    ql.set_quant_state(True, True)
    return ql


@pytest.mark.parametrize("tag", ["conv3", "conv3s2", "conv1"])
@pytest.mark.parametrize("mode", ["g1", "g1u", "kwise", "rowwise"])
def test_quant_layer_conv_api(ops_golden, tag, mode):
    c = ops_golden[f"{tag}_{mode}"]
    bsz, ci, hw, co, k, s, wbits = c["shape"]
    x = torch.randn(bsz, ci, hw, hw, generator=torch.Generator().manual_seed(c["seed"]))
    ql = make_layer(nn.Conv2d(ci, co, k, s, k // 2), c, wbits, c["grouped"])
    y = ql(x.to(DEV))
    assert y.shape == c["out"].shape and y.dtype == torch.float32
    # weight quantizer initialised by per-channel MINMAX exactly as the reference's first forward does
    assert torch.equal(ql.wqtizer.delta.cpu(), c["wdelta"]) and torch.equal(ql.wqtizer.zero_point.cpu(), c["wzp"])
    # scalar / row-wise scales: integer operands, exact accumulation -> fp32-level agreement;
    # K-wise scales are folded into the fp16 operand (2^-11 relative rounding)
    tol = 2e-3 if mode == "kwise" else 2e-5
    assert rel(y, c["out"]) < tol, rel(y, c["out"])


@pytest.mark.parametrize("tag", ["lin_w4", "lin_w8"])
@pytest.mark.parametrize("mode", ["g1", "kwise", "rowwise"])
def test_quant_layer_linear_api(ops_golden, tag, mode):
    c = ops_golden[f"{tag}_{mode}"]
    x = torch.randn(2, 50, 96, generator=torch.Generator().manual_seed(c["seed"]))
    ql = make_layer(nn.Linear(96, 80), c, 4 if tag == "lin_w4" else 8, True)
    y = ql(x.to(DEV))
    tol = 2e-3 if mode == "kwise" else 2e-5
    assert rel(y, c["out"]) < tol, rel(y, c["out"])
    # quantization off -> plain fp16-operand GEMM on the original weights
    ql.set_quant_state(False, False)
    y = ql(x.to(DEV))
    ref = torch.nn.functional.linear(x, c["weight"], c["bias"])
    assert rel(y, ref) < 2e-3


def test_quantizer_modules_api(ops_golden):
    from quant.quant_layer import UniformAffineQuantizer, Scaler
    from quant.quant_layer_text import T2ILogQuantizer
    for bits in (8, 6):
        for layout in ("scalar", "in", "out"):
            c = ops_golden[f"uaq_a{bits}_{layout}"]
            x = torch.randn(2, 48, 40, generator=torch.Generator().manual_seed(c["seed"])) * 2
            q = UniformAffineQuantizer(bits=bits, channel_wise=False, scaler=Scaler.MINMAX)
            q.delta, q.zero_point, q.init = c["delta"].to(DEV), c["zp"].to(DEV), True
            assert torch.equal(q(x.to(DEV)).cpu(), c["out"])
    c = ops_golden["t2i_log_rt1"]
    p = torch.softmax(torch.randn(2, 4, 33, 77, generator=torch.Generator().manual_seed(c["seed"])) * 3, dim=-1)
    q = T2ILogQuantizer(bits=8, real_time=True)
    out = q(p.to(DEV)).cpu()
    assert (out != c["out"]).float().mean().item() < 1e-4
    with pytest.raises(RuntimeError):
        UniformAffineQuantizer(bits=8)(torch.zeros(4))          # CPU tensors: no fallback path


@pytest.mark.parametrize("tag,log,rt,sp", [("uni", False, False, False), ("log_rt", True, True, False),
                                           ("log_rt_sp", True, True, True), ("log_static", True, False, False)])
def test_transformer_block_attention_api(ops_golden, tag, log, rt, sp):
    """QuantBasicTransformerBlock built like the reference's fixture: plain block -> QuantModel-style
    wrapping with quantization of the inner linears off, attention quantizers on."""
    from quant.quant_layer import QuantLayer, Scaler
    from quant.quant_block import QuantBasicTransformerBlock
    from dgq_b200.unet import sd as graph
    c = ops_golden[f"attn_{tag}"]
    g = torch.Generator().manual_seed(c["seed"])
    x = torch.randn(2, 36, 64, generator=g)
    ctx = torch.randn(2, 77, 768, generator=g)
    blk = graph.BasicTransformerBlock(64)
    aq = {"bits": 8, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True}
    wq = {"bits": 8, "channel_wise": True, "scaler": Scaler.MINMAX}
    for mod in list(blk.modules()):                              # what QuantModel.quant_module does
        for name, child in list(mod.named_children()):
            if isinstance(child, nn.Linear):
                setattr(mod, name, QuantLayer(child, wq, aq))
    qb = QuantBasicTransformerBlock(blk, aq, {"softmax_a_bit": 8, "t2i_log_quant": log, "t2i_real_time": rt,
                                              "t2i_start_peak": sp, "log_max_1": False})
    state = {}
    for k, v in c["state"].items():
        if ".to_" in k or k.startswith("ff."):
            k = k.replace(".weight", ".w").replace(".bias", ".b")
        state[k] = v
    missing = qb.load_state_dict(state, strict=False)
    assert not [k for k in missing.missing_keys if k.endswith((".w", ".b", ".weight", ".bias"))]
    qb = qb.to(DEV)
    for m in qb.modules():                                       # original_w is cloned at construction
        if isinstance(m, QuantLayer):
            m.original_w = m.w.data.clone(); m.original_b = None if m.b is None else m.b.data.clone()
    for an in ("attn1", "attn2"):
        a = getattr(qb, an)
        a.use_aq = True
        for qn in ("aqtizer_q", "aqtizer_k", "aqtizer_v"):
            d, z = c["params"][f"{an}.{qn}"]
            q = getattr(a, qn)
            q.delta, q.zero_point, q.init = d.to(DEV), z.to(DEV), True
        if not log:
            a.aqtizer_w.delta, a.aqtizer_w.zero_point, a.aqtizer_w.init = torch.tensor(1 / 255., device=DEV), torch.tensor(0., device=DEV), True
        elif not rt:
            a.aqtizer_w.delta, a.aqtizer_w.init = torch.tensor(0.41, device=DEV), True
    o1 = qb.attn1(x.to(DEV))
    o2 = qb.attn2(x.to(DEV), encoder_hidden_states=ctx.to(DEV))
    for o, gold in ((o1, c["out1"]), (o2, c["out2"])):
        cos = torch.nn.functional.cosine_similarity(o.cpu().flatten(), gold.flatten(), dim=0).item()
        l2 = ((o.cpu() - gold).norm() / gold.norm()).item()
        assert cos > 0.9995 and l2 < 3e-2, (cos, l2)
