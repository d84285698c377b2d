"""ctypes binding of include/dgq_b200.h.  Fails loudly when the CUDA library is missing: there is
no CPU or eager fallback behind these calls."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_C", "libdgq_b200.so")

SYMBOLS = [
    "dgq_version", "dgq_fake_quant_f32", "dgq_t2i_log_quant_f32", "dgq_max_f32", "dgq_pack_weight", "dgq_unpack_weight",
    "dgq_act_producer", "dgq_gn_stats", "dgq_ln_quant", "dgq_row_quant", "dgq_geglu_quant",
    "dgq_gemm_f16", "dgq_gemm_i8", "dgq_weight_to_i8", "dgq_conv_oob_colsum", "dgq_qkv_pack", "dgq_attention", "dgq_timestep_embedding", "dgq_nchw_to_nhwc",
    "dgq_nhwc_to_nchw", "dgq_silu", "dgq_add", "dgq_softmax_rows", "dgq_sampler_step",
]

Q_NONE, Q_SCALAR, Q_KWISE, Q_ROWWISE = 0, 1, 2, 3
MAP_NONE, MAP_UNIFORM, MAP_LOG2 = 0, 1, 2
EPI_PLAIN, EPI_GEGLU, EPI_QKV = 0, 1, 2


class QuantT(C.Structure):
    _fields_ = [("delta", C.c_void_p), ("zp", C.c_void_p), ("inv_delta", C.c_void_p), ("mode", C.c_int), ("period", C.c_int),
                ("qmax", C.c_float), ("emit_int", C.c_int)]


class ProducerT(C.Structure):
    _fields_ = [("src0", C.c_void_p), ("src1", C.c_void_p), ("c0", C.c_int), ("c1", C.c_int),
                ("src_is_f32", C.c_int), ("batch", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("upsample", C.c_int), ("ksize", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
                ("gn_mean", C.c_void_p), ("gn_rstd", C.c_void_p), ("gn_gamma", C.c_void_p),
                ("gn_beta", C.c_void_p), ("act", C.c_int), ("q", QuantT), ("pad_quantized", C.c_int),
                ("out", C.c_void_p), ("ldo", C.c_int), ("codes", C.c_void_p)]


class GemmT(C.Structure):
    _fields_ = [("a", C.c_void_p), ("lda", C.c_int), ("b", C.c_void_p), ("ldb", C.c_int),
                ("m", C.c_int), ("n", C.c_int), ("k", C.c_int), ("scale", C.c_void_p),
                ("row_scale", C.c_void_p), ("row_period", C.c_int), ("bias", C.c_void_p), ("temb", C.c_void_p), ("rows_per_batch", C.c_int),
                ("ld_temb", C.c_int), ("resid", C.c_void_p), ("ld_resid", C.c_int),
                ("out", C.c_void_p), ("ldc", C.c_int), ("out_f32", C.c_void_p), ("ep_is_f32", C.c_int),
                ("epi", C.c_int), ("q2", QuantT), ("heads", C.c_int), ("d", C.c_int), ("dp", C.c_int),
                ("tokens", C.c_int), ("tp", C.c_int), ("transpose", C.c_int), ("skip_first", C.c_int),
                ("kfold", C.c_void_p), ("k_split", C.c_int),
                ("colsum", C.c_void_p), ("b_off", C.c_void_p), ("row_zp", C.c_void_p),
                ("conv_b", C.c_int), ("conv_h", C.c_int), ("conv_w", C.c_int), ("conv_c", C.c_int),
                ("conv_csoob", C.c_void_p), ("conv_ldoob", C.c_int)]


class AttnT(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("vt", C.c_void_p), ("b", C.c_int),
                ("heads", C.c_int), ("t", C.c_int), ("s", C.c_int), ("sp", C.c_int), ("d", C.c_int),
                ("dp", C.c_int), ("scale", C.c_float), ("map_mode", C.c_int), ("real_time", C.c_int),
                ("start_peak", C.c_int), ("delta", C.c_void_p), ("qmax", C.c_float),
                ("row_max", C.c_void_p), ("row_sum", C.c_void_p), ("gmax", C.c_void_p),
                ("out", C.c_void_p), ("ldo", C.c_int), ("out_is_f32", C.c_int), ("codes", C.c_void_p),
                ("out_q", QuantT), ("q_scale", C.c_void_p), ("q_scale_period", C.c_int), ("k_split", C.c_int)]


class SamplerStepT(C.Structure):
    _fields_ = [("unet_out", C.c_void_p), ("n", C.c_int64), ("guidance", C.c_float), ("use_cfg", C.c_int),
                ("eps_store", C.c_void_p), ("x", C.c_void_p), ("cx", C.c_float), ("c_eps", C.c_float),
                ("hist", C.c_void_p * 4), ("c_hist", C.c_float * 4), ("noise", C.c_void_p), ("c_noise", C.c_float),
                ("out", C.c_void_p), ("model_in", C.c_void_p), ("in_scale", C.c_float), ("dup", C.c_int)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"dgq_b200: CUDA library {LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no fallback path.")
        l = C.CDLL(LIB_PATH)
        vp, i, i64, f = C.c_void_p, C.c_int, C.c_int64, C.c_float
        sig = {
            "dgq_version": [],
            "dgq_fake_quant_f32": [vp, i64, vp, vp, i, i64, f, vp, vp, vp],
            "dgq_t2i_log_quant_f32": [vp, i64, vp, f, vp, vp, vp],
            "dgq_max_f32": [vp, i64, vp, vp, vp],
            "dgq_pack_weight": [vp, vp, vp, vp, i, i, i, i, i, f, i, vp, vp, vp, vp],
            "dgq_unpack_weight": [vp, i, vp, i, i, i, i, i, vp, vp],
            "dgq_act_producer": [C.POINTER(ProducerT), vp],
            "dgq_gn_stats": [vp, vp, i, i, i, i, i, f, vp, vp, vp, vp],
            "dgq_ln_quant": [vp, i, i, i, vp, vp, f, i, C.POINTER(QuantT), C.POINTER(vp), vp],
            "dgq_row_quant": [vp, i, i, i, i, C.POINTER(QuantT), C.POINTER(vp), C.POINTER(vp), vp],
            "dgq_geglu_quant": [vp, i, i, i, QuantT, vp, vp],
            "dgq_gemm_f16": [C.POINTER(GemmT), vp],
            "dgq_gemm_i8": [C.POINTER(GemmT), vp],
            "dgq_weight_to_i8": [vp, vp, i, i, i, f, vp, vp, vp, vp],
            "dgq_conv_oob_colsum": [vp, i, i, vp, vp],
            "dgq_qkv_pack": [vp, i, i, i, i, i, i, i, i, i, i, QuantT, vp, i, vp, vp],
            "dgq_attention": [C.POINTER(AttnT), vp],
            "dgq_timestep_embedding": [vp, i, i, vp, vp, i, vp],
            "dgq_nchw_to_nhwc": [vp, i, i, i, i, vp, i, vp],
            "dgq_nhwc_to_nchw": [vp, i, i, i, i, i, vp, vp],
            "dgq_silu": [vp, i, i64, vp, vp],
            "dgq_softmax_rows": [vp, i64, i, i64, f, vp, i64, vp],
            "dgq_add": [vp, vp, i, i64, vp, vp],
            "dgq_sampler_step": [C.POINTER(SamplerStepT), vp],
        }
        for name, args in sig.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = l
    return _lib


class DgqError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = "invalid argument" if rc == -1 else f"cudaError {rc}"
        raise DgqError(f"{what}: {msg}")
