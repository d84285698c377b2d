"""kind::i8 qGEMM path (dgq_gemm_i8: u8 activation codes x s8 weight codes, s32 accumulate, integer zero-point
corrections) against exact integer arithmetic and against the kind::f16 exact-integer path it replaces for
scalar / row-wise activation scales (reference quant/quant_layer.py:295-299, 626-661)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scales(g, n, level):
    lo = -(torch.rand(n, generator=g) * 3 + 1)
    hi = torch.rand(n, generator=g) * 3 + 1
    d = (hi - lo) / (level - 1)
    return d, torch.round(-lo / d)


@pytest.mark.parametrize("m,n,k", [(128, 256, 128), (300, 200, 320), (1000, 1280, 1280), (4096, 320, 2880), (77, 640, 2048),
                                   (64, 1280, 11520), (1, 1280, 1280), (256, 1288, 2304)])   # the last three: 64-column tiles, 6-deep ring
@pytest.mark.parametrize("wbits", [4, 8])
@pytest.mark.parametrize("rowwise", [False, True])
def test_gemm_i8_exact(m, n, k, wbits, rowwise):
    """C = dA[m] dW[n] sum_k (a - za[m]) (w - wz[n]) + bias, against the same sum in int64."""
    from dgq_b200 import ops
    g = torch.Generator().manual_seed(m + n + k + wbits)
    a = torch.randint(0, 256, (m, k), generator=g, dtype=torch.int32)
    wlevel = 2 ** wbits
    wc = torch.randint(0, wlevel, (n, k), generator=g, dtype=torch.int32)
    wz = torch.randint(0, wlevel, (n,), generator=g).float()
    wd = torch.rand(n, generator=g) * 0.01 + 0.001
    bias = torch.randn(n, generator=g)
    period = 50 if rowwise else 1
    ad, az = _scales(g, period, 256)
    if rowwise:
        az[0], az[1] = -20.0, 300.0          # zero points outside [0, 255] occur in grouped mode
    rows = torch.arange(m) % period
    ref = ((a.long() - az[rows].long()[:, None]) @ (wc.long() - wz.long()[:, None]).t()).double()
    ref = ref * ad[rows].double()[:, None] * wd.double()[None, :] + bias.double()[None, :]
    n_pad = (n + 7) // 8 * 8
    codes = torch.zeros(n_pad, k, dtype=torch.uint8)
    codes[:n] = wc.to(torch.uint8)
    op, colsum, b_off = ops.weight_to_i8(codes.to(DEV), wz.to(DEV), n, float(wlevel - 1))
    assert (b_off is None) == (wbits == 4)
    scale = torch.zeros(n_pad); scale[:n] = wd
    bb = torch.zeros(n_pad); bb[:n] = bias
    y = ops.gemm(a.to(torch.uint8).to(DEV), op, n_pad, scale=scale.to(DEV), bias=bb.to(DEV), want_f32=True,
                 row_scale=ad.to(DEV), row_period=period, row_zp=az.to(DEV), colsum=colsum, b_off=b_off)
    err = ((y[:, :n].double().cpu() - ref).abs().max() / ref.abs().max()).item()
    assert err < 2e-6, err


def _layer(layer, wbits, abits, d, z, grouped):
    from quant.quant_layer import QuantLayer, Scaler
    ql = QuantLayer(layer, {"bits": wbits, "channel_wise": True, "scaler": Scaler.MINMAX},
                    {"bits": abits, "channel_wise": False, "scaler": Scaler.MINMAX, "leaf_param": True}).to(DEV)
    ql.aqtizer.delta, ql.aqtizer.zero_point, ql.aqtizer.init = d.to(DEV), z.to(DEV), True
    ql.use_group_num = grouped
    ql.set_quant_state(True, True)
    return ql


@pytest.mark.parametrize("wbits,abits", [(8, 8), (8, 6), (4, 8)])
@pytest.mark.parametrize("kind", ["conv3", "conv3_s2", "conv3_rowwise", "conv1", "linear", "linear_rowwise"])
def test_layer_i8_matches_f16_exact_path(wbits, abits, kind):
    """QuantLayer.forward through the i8 path == the kind::f16 exact-integer path (both accumulate the same integers
    exactly; only the order of the fp32 scale multiplications differs) and == the oracle."""
    from dgq_b200 import engine, ops
    from oracle import dgq_oracle as O
    g = torch.Generator().manual_seed(7 + wbits + abits)
    torch.manual_seed(3)
    level = 2 ** abits
    if kind.startswith("conv"):
        k = 1 if kind == "conv1" else 3
        s = 2 if kind == "conv3_s2" else 1
        ci, co, hw, bsz = (128, 96, 24, 2)
        layer = nn.Conv2d(ci, co, k, s, k // 2)
        x = torch.randn(bsz, ci, hw, hw, generator=g) * 1.5
        ho = (hw + 2 * (k // 2) - k) // s + 1
        if kind == "conv3_rowwise":
            d, z = _scales(g, ho * ho, level)
            d, z, grouped = d.view(1, 1, -1), z.view(1, 1, -1), True
        else:
            d, z = _scales(g, 1, level)
            d, z, grouped = d[0], z[0], False
    else:
        layer = nn.Linear(320, 200)
        x = torch.randn(3, 50, 320, generator=g) * 1.5
        if kind == "linear_rowwise":
            d, z = _scales(g, 50, level)
            d, z, grouped = d.view(1, -1, 1), z.view(1, -1, 1), True
        else:
            d, z = _scales(g, 1, level)
            d, z, grouped = d[0], z[0], False
    w_cpu, b_cpu = layer.weight.detach().clone(), layer.bias.detach().clone()
    ql = _layer(layer, wbits, abits, d, z, grouped)
    q = ql.act_qparam(torch.device(DEV))
    assert ql.i8_ok(q)
    n0 = ops.LAUNCHES
    engine.USE_I8 = True
    try:
        y8 = ql(x.to(DEV))
        engine.USE_I8 = False
        y16 = ql(x.to(DEV))
    finally:
        engine.USE_I8 = True
    assert ops.LAUNCHES > n0
    wd, wz = O.channel_minmax_scale(w_cpu, 2 ** wbits)
    sd = {"l.w": w_cpu, "l.b": b_cpu, "l.wqtizer.delta": wd, "l.wqtizer.zero_point": wz}
    cfg = O.QConfig(wbits=wbits, abits=abits, group_convs={"l"} if (grouped and kind.startswith("conv")) else set())
    ref = O.quant_layer(x, sd, {"l.aqtizer.delta": d, "l.aqtizer.zero_point": z}, "l", cfg,
                        stride=2 if kind == "conv3_s2" else 1, padding=1 if kind.startswith("conv3") else 0)
    scale = ref.abs().max()
    assert ((y8.cpu() - y16.cpu()).abs().max() / scale).item() < 1e-6
    assert ((y8.cpu() - ref).abs().max() / scale).item() < 2e-5


@pytest.mark.parametrize("mode", ["scalar", "rowwise"])
def test_producers_emit_u8_codes(mode):
    """emit_int = 2 writes exactly the integer codes the verification output (want_codes) reports."""
    from dgq_b200 import ops
    g = torch.Generator().manual_seed(11)
    # rows: LayerNorm + quantize (tiled kernel, c = 640) and the plain row quantizer (c = 96)
    for m, c in ((400, 640), (77, 96)):
        x = (torch.randn(m, c, generator=g) * 2).to(DEV)
        d, z = _scales(g, 50 if mode == "rowwise" else 1, 256)
        qp = ops.qparam_from_ckpt(d.view(1, -1, 1) if mode == "rowwise" else d[0], z.view(1, -1, 1) if mode == "rowwise" else z[0],
                                  255.0, DEV)
        u8 = ops.row_quant(x, [qp], emit_int=2)[0]
        _, codes = ops.row_quant(x, [qp], want_codes=True)
        assert u8.dtype == torch.uint8 and torch.equal(u8, codes[0])
        if c == 640:
            gamma, beta = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
            ln8 = ops.ln_quant(x, gamma, beta, 1e-5, [qp], emit_int=2)[0]
            lni = ops.ln_quant(x, gamma, beta, 1e-5, [qp], emit_int=1)[0]
            zrow = qp.zp[torch.arange(m, device=DEV) % qp.period][:, None]
            assert torch.equal(ln8.float(), lni.float() + zrow)
    # im2col: tiled kernel (c % 64 == 0, stride 1) and the generic one (stride 2)
    for c, stride in ((64, 1), (40, 2)):
        b, h = 2, 16
        src = (torch.randn(b, h, h, c, generator=g) * 2).to(DEV)
        ho = (h + 2 - 3) // stride + 1
        d, z = _scales(g, ho * ho if mode == "rowwise" else 1, 256)
        qp = ops.qparam_from_ckpt(d.view(1, 1, -1) if mode == "rowwise" else d[0], z.view(1, 1, -1) if mode == "rowwise" else z[0],
                                  255.0, DEV, conv=True)
        padq = mode == "rowwise"
        u8 = ops.act_producer(src, batch=b, h=h, w=h, ksize=3, stride=stride, q=qp, pad_quantized=padq, emit_int=2)
        _, codes = ops.act_producer(src, batch=b, h=h, w=h, ksize=3, stride=stride, q=qp, pad_quantized=padq,
                                    want_codes=True)
        if padq:
            assert torch.equal(u8, codes)
        else:        # exact-zero padding: the verification output marks those taps 0, the operand holds the code zp
            fi = ops.act_producer(src, batch=b, h=h, w=h, ksize=3, stride=stride, q=qp, pad_quantized=False, emit_int=1)
            assert torch.equal(u8.float(), fi.float() + qp.zp[0])


@pytest.mark.parametrize("bsz,ci,co,hw", [(1, 64, 96, 8), (2, 64, 96, 8), (3, 64, 40, 8), (2, 320, 320, 16), (1, 128, 64, 32),
                                          (2, 192, 640, 64), (1, 320, 320, 64)])
@pytest.mark.parametrize("wbits,abits", [(8, 8), (8, 6), (4, 8)])
def test_implicit_conv_matches_im2col_and_oracle(bsz, ci, co, hw, wbits, abits):
    """3x3 / stride 1 / pad 1 conv with a per-tensor activation scale (reference quant_layer.py:659: F.conv2d on
    x_hat with EXACT-zero padding): the implicit GEMM (NHWC codes gathered by 4-D TMA, border-class zero-point
    correction) == the im2col producer + GEMM == the oracle.  Patch shapes 8x8x2 / 16x8 are all exercised."""
    from dgq_b200 import engine, ops
    from oracle import dgq_oracle as O
    g = torch.Generator().manual_seed(bsz + ci + hw + wbits)
    torch.manual_seed(5)
    layer = nn.Conv2d(ci, co, 3, 1, 1)
    x = torch.randn(bsz, ci, hw, hw, generator=g) * 1.5
    d, z = _scales(g, 1, 2 ** abits)
    w_cpu, b_cpu = layer.weight.detach().clone(), layer.bias.detach().clone()
    ql = _layer(layer, wbits, abits, d[0], z[0], False)
    q = ql.act_qparam(torch.device(DEV))
    assert engine._implicit_ok(ql, q, hw, hw, 3, 1)
    try:
        engine.IMPLICIT_CONV = True
        ql(x.to(DEV))                   # first call packs the weights (three one-off launches)
        n0 = ops.LAUNCHES
        y_imp = ql(x.to(DEV))
        n_imp = ops.LAUNCHES - n0
        engine.IMPLICIT_CONV = False
        y_col = ql(x.to(DEV))
    finally:
        engine.IMPLICIT_CONV = True
    assert n_imp == 2           # one NHWC quantize launch + one GEMM
    wd, wz = O.channel_minmax_scale(w_cpu, 2 ** wbits)
    sd = {"l.w": w_cpu, "l.b": b_cpu, "l.wqtizer.delta": wd, "l.wqtizer.zero_point": wz}
    ref = O.quant_layer(x, sd, {"l.aqtizer.delta": d[0], "l.aqtizer.zero_point": z[0]}, "l",
                        O.QConfig(wbits=wbits, abits=abits), padding=1)
    scale = ref.abs().max()
    assert ((y_imp.cpu() - y_col.cpu()).abs().max() / scale).item() < 1e-6
    assert ((y_imp.cpu() - ref).abs().max() / scale).item() < 2e-5
