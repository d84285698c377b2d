"""Compiled checkpoints: compile a loaded QuantModel, load it back without any fp32 master
weights, and require BIT-IDENTICAL UNet outputs (same operands, same kernels); plus the unpack kernel alone and the
corruption check."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("shape", [(320, 320, 3, 3), (640, 1280), (4, 320, 3, 3), (1280, 2816)])
def test_unpack_weight_matches_pack(bits, shape):
    from dgq_b200 import ops
    from dgq_b200.quant.quant_layer import channel_minmax
    g = torch.Generator().manual_seed(sum(shape) + bits)
    w = (torch.randn(*shape, generator=g) * 0.1).to(DEV)
    d, z = channel_minmax(w, 2 ** bits)
    n = shape[0]
    n_pad = (n + 7) // 8 * 8
    operand, codes, packed = ops.pack_weight(w, d, z, None, float(2 ** bits - 1), True, n_pad=n_pad, want_codes=True,
                                             want_packed4=bits == 4)
    ci = shape[1]
    taps = shape[2] * shape[3] if len(shape) == 4 else 1
    back = ops.unpack_weight(packed if bits == 4 else codes, bits, z.reshape(-1), n, ci, taps, (ci + 7) // 8 * 8, n_pad)
    assert torch.equal(back, operand)


@pytest.mark.parametrize("model_type,kw", [
    ("sd", dict(wbits=4, abits=8, group_num=8, n_steps=2)),
    ("sd", dict(wbits=8, abits=8, group_num=1, n_steps=1, log_quant=False, real_time=False, start_peak=False)),
])
def test_compiled_roundtrip_bit_identical(model_type, kw, tmp_path):
    from dgq_b200 import compiled, synthetic
    from dgq_b200.quant.quant_layer import QuantLayer
    qnn = synthetic.make_qmodel(model_type, device=DEV, seed=3, **kw)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 32, 32, generator=g).to(DEV)
    ctx = torch.randn(2, 77, 768, generator=g).to(DEV)
    ts = [torch.tensor([981.0]), torch.tensor([401.0])][: kw["n_steps"]]
    with torch.no_grad():
        ref = [qnn(x, t, ctx)[0].clone() for t in ts]
    path = os.path.join(tmp_path, "unet.dgqb")
    header = compiled.compile_checkpoint(qnn, path)
    size = os.path.getsize(path)
    n_w = sum(m.w.numel() for m in qnn.modules() if isinstance(m, QuantLayer))
    print(f"{model_type} W{kw['wbits']}: {size / 2**20:.0f} MiB on disk for {n_w / 1e6:.0f} M weights "
          f"({8 * size / n_w:.2f} bits/weight incl. scales and activation tables)")
    assert size < n_w * (kw["wbits"] / 8 + 0.2)
    del qnn
    torch.cuda.empty_cache()
    q2 = compiled.load_compiled(path, DEV)
    assert all(m.w.is_meta for m in q2.modules() if isinstance(m, QuantLayer))   # no master weights
    with torch.no_grad():
        out = [q2(x, t, ctx)[0] for t in ts]
    for a, b in zip(out, ref):
        assert torch.equal(a, b), (a - b).abs().max().item()
    assert compiled.read_header(path)[0]["sha256"] == header["sha256"]
    # a flipped payload byte is detected
    raw = bytearray(open(path, "rb").read())
    raw[len(raw) // 2] ^= 0x40          # inside a weight-code tensor (the tail of the file is alignment padding)
    bad = os.path.join(tmp_path, "bad.dgqb")
    open(bad, "wb").write(raw)
    with pytest.raises(ValueError):
        compiled.load_compiled(bad, DEV)
    with pytest.raises(RuntimeError):
        compiled.load_compiled(path, "cpu")
