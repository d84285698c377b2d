// A-operand producers: everything between two GEMMs of the quantized UNet, fused with the
// activation quantizer of the consuming QuantLayer.  All are HBM-bound, 16-byte vectorised,
// one pass over their input (taps of a 3x3 window re-read through L1/L2).
//
//   dgq_act_producer  [concat] -> [nearest x2] -> [GroupNorm] -> [SiLU] -> im2col -> quantize
//   dgq_gn_stats      GroupNorm(32) mean / rstd (deterministic two-stage reduction)
//   dgq_ln_quant      LayerNorm + up to 3 quantizers
//   dgq_row_quant     quantizer only (fp32 or fp16 rows)
//   dgq_geglu_quant   x1 * gelu(x2) + quantizer
//   dgq_qkv_pack      head split (+transpose for V) + quantizer -> attention operands
//
// Reference semantics: quant/quant_layer.py:295-299 (quantizer), :630-641 (unfold then quantize:
// zero padding IS quantised on that path), quant/quant_block.py:98-119.
#include "common.cuh"

namespace dgq {

struct QuantDev {
  const float* delta;
  const float* zp;
  const float* inv;   // 1/delta (correctly rounded) or nullptr
  int mode;
  int period;
  float qmax;
  int emit_int;
};

static QuantDev to_dev(const dgq_quant_t& q) {
  return QuantDev{q.delta, q.zp, q.inv_delta, q.mode, q.period, q.qmax, q.emit_int};
}

// zero point of row `row` for the SCALAR / ROWWISE quantizers (the u8-code output adds it back)
__device__ __forceinline__ float quant_zp(const QuantDev& q, int row) {
  if (q.mode == DGQ_Q_SCALAR) return __ldg(q.zp);
  if (q.mode == DGQ_Q_ROWWISE) return __ldg(q.zp + row % q.period);
  return 0.0f;
}
// store 8 consecutive operand values of one row at element index `idx`: fp16 (16 bytes), or -- emit_int == 2 -- the
// u8 codes (8 bytes) of the kind::i8 GEMM, rebuilt from the integer (code - zp) the quantizer returned
__device__ __forceinline__ void store8(void* out, size_t idx, const float (&v)[8], int emit_int, float zp) {
  if (emit_int == 2) {
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i >> 2] |= static_cast<uint32_t>(__float2int_rn(v[i] + zp)) << (8 * (i & 3));
    *reinterpret_cast<uint2*>(static_cast<uint8_t*>(out) + idx) = make_uint2(w[0], w[1]);
  } else {
    *reinterpret_cast<uint4*>(static_cast<__half*>(out) + idx) = pack8(v);
  }
}

// quantize 8 consecutive K positions k0..k0+7 of row `row` in place (no code output): the lean path
// of every fused producer.  KWISE reads (delta, 1/delta, zp) vectors, SCALAR / ROWWISE one triple.
__device__ __forceinline__ void quant8_lean(const QuantDev& q, float (&v)[8], int k0, int row) {
  if (q.mode == DGQ_Q_NONE) return;
  if (q.mode == DGQ_Q_KWISE) {
    float d[8], z[8], inv[8];
    ldg8(q.delta + k0, d);
    ldg8(q.zp + k0, z);
    if (q.inv != nullptr) {
      ldg8(q.inv + k0, inv);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) inv[i] = rcp_rn_slow(d[i]);
    }
    float lo[8], hi[8];
    uaq_bounds<8>(z, q.qmax, lo, hi);
    uaq_lean_lh<false, 8>(v, d, inv, lo, hi);   // emit_int is not defined for K-wise scales
  } else {
    const int j = q.mode == DGQ_Q_ROWWISE ? row % q.period : 0;
    const float dd = __ldg(q.delta + j), zz = __ldg(q.zp + j);
    const float ii = q.inv != nullptr ? __ldg(q.inv + j) : rcp_rn_slow(dd);
    if (q.emit_int) uaq_lean1_lh<true, 8>(v, dd, ii, -zz, __fsub_rn(q.qmax, zz));
    else uaq_lean1_lh<false, 8>(v, dd, ii, -zz, __fsub_rn(q.qmax, zz));
  }
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
__device__ __forceinline__ void from_f(float v, float* o) { *o = v; }
__device__ __forceinline__ void from_f(float v, __half* o) { *o = __float2half_rn(v); }

// quantize 8 consecutive K positions k0..k0+7 of row `row`; returns de-quantised values in v
__device__ __forceinline__ void quant8(const QuantDev& q, float (&v)[8], int k0, int row, uint8_t* codes8) {
  if (q.mode == DGQ_Q_NONE) return;
  float d[8], z[8];
  if (q.mode == DGQ_Q_KWISE) {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(q.delta + k0));
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(q.delta + k0 + 4));
    const float4 z0 = __ldg(reinterpret_cast<const float4*>(q.zp + k0));
    const float4 z1 = __ldg(reinterpret_cast<const float4*>(q.zp + k0 + 4));
    d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
    z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
  } else {
    const int j = q.mode == DGQ_Q_ROWWISE ? row % q.period : 0;
    const float dd = __ldg(q.delta + j), zz = __ldg(q.zp + j);
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = dd; z[i] = zz; }
  }
  uint32_t lo = 0, hi = 0;
  float inv[8], cd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) inv[i] = __frcp_rn(d[i]);
  uaq_codes_rcp<8>(v, d, inv, z, q.qmax, cd);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float c = cd[i];
    v[i] = q.emit_int ? __fsub_rn(c, z[i]) : uaq_dequant(c, d[i], z[i]);
    if (i < 4) lo |= static_cast<uint32_t>(c) << (8 * i);
    else hi |= static_cast<uint32_t>(c) << (8 * (i - 4));
  }
  if (codes8 != nullptr) *reinterpret_cast<uint2*>(codes8) = make_uint2(lo, hi);
}

// ------------------------------------------------------------------------------------------
struct ProducerDev {
  const void* src0;
  const void* src1;
  int c0, c1;
  int batch, h, w, hs, ws, ho, wo;
  int upsample, ksize, stride, pad;
  const float* gn_mean;
  const float* gn_rstd;
  const float* gn_gamma;
  const float* gn_beta;
  int act;
  QuantDev q;
  int pad_quantized;
  __half* out;
  int ldo;
  uint8_t* codes;
};

template <typename TIn>
__global__ void __launch_bounds__(256) act_producer_kernel(const ProducerDev p) {
  const int C = p.c0 + p.c1;
  const int K = p.ksize * p.ksize * C;
  const int kvec = p.ldo >> 3;
  const int64_t total = static_cast<int64_t>(p.batch) * p.ho * p.wo * kvec;
  const int cpg = C >> 5;  // channels per GroupNorm group
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(idx / kvec);
    const int k0 = static_cast<int>(idx % kvec) << 3;
    const size_t dst = static_cast<size_t>(m) * p.ldo + k0;
    if (k0 >= K) {  // zero padding of the K tail (ldo > K)
      if (p.q.emit_int == 2) *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(p.out) + dst) = make_uint2(0, 0);
      else *reinterpret_cast<uint4*>(p.out + dst) = make_uint4(0, 0, 0, 0);
      continue;
    }
    const int tap = k0 / C, c = k0 % C;
    const int ox = m % p.wo;
    const int oy = (m / p.wo) % p.ho;
    const int b = m / (p.wo * p.ho);
    const int iy = oy * p.stride - p.pad + tap / p.ksize;
    const int ix = ox * p.stride - p.pad + tap % p.ksize;
    float v[8];
    const bool inside = iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
    if (inside) {
      const int sy = p.upsample ? (iy >> 1) : iy, sx = p.upsample ? (ix >> 1) : ix;
      const size_t pix = (static_cast<size_t>(b) * p.hs + sy) * p.ws + sx;
      if (c < p.c0) load8(static_cast<const TIn*>(p.src0) + pix * p.c0 + c, v);
      else load8(static_cast<const TIn*>(p.src1) + pix * p.c1 + (c - p.c0), v);
      if (p.gn_mean != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int g = (c + i) / cpg;
          const float mean = __ldg(p.gn_mean + b * 32 + g), rstd = __ldg(p.gn_rstd + b * 32 + g);
          v[i] = (v[i] - mean) * rstd * __ldg(p.gn_gamma + c + i) + __ldg(p.gn_beta + c + i);
        }
      }
      if (p.act == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = silu_f(v[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.0f;
    }
    uint8_t* cdst = p.codes != nullptr ? p.codes + static_cast<size_t>(m) * K + k0 : nullptr;
    if (inside || p.pad_quantized) quant8(p.q, v, k0, m, cdst);
    else if (cdst != nullptr) *reinterpret_cast<uint2*>(cdst) = make_uint2(0, 0);
    // an un-quantized padding tap is an exact 0 = (code - zp) 0, i.e. the code zp in the u8 operand
    store8(p.out, dst, v, p.q.emit_int, p.q.emit_int == 2 ? quant_zp(p.q, m) : 0.0f);
  }
}

// ------------------------------------------------------------------------------------------
// Tiled im2col producer for the hot shapes (ksize 1 or 3, stride 1, channels a multiple of 64).
// A CTA owns an 8 x 16 output-pixel tile x 64 channels: phase 1 loads the (8+2) x (16+2) input
// patch ONCE (concat / nearest-x2 / GroupNorm / SiLU applied once per input element instead of
// once per tap), phase 2 emits the 9 taps -- each with its own (delta, zp) when the scales are
// K-wise -- as 128-byte row segments of the A operand.  The quantizer runs on the reciprocal fast
// path (uaq_codes_rcp, bit-identical codes).
constexpr int kTileH = 8, kTileW = 16, kTileC = 64;

template <typename TIn, int KS, int QMODE, bool kCodes>
__global__ void __launch_bounds__(256) conv_producer_kernel(const ProducerDev p, int tiles_x, int tiles_y) {
  constexpr int PH = kTileH + KS - 1, PW = kTileW + KS - 1;
  // per pixel: two planes of 8 float4 -- plane q holds channels 8 s + 4 q .. + 3 of every 8-channel slot s, so the
  // phase-2 reads (lane = slot, one float4 per plane) are 128 contiguous bytes per quarter-warp: no bank conflict
  __shared__ __align__(16) float patch[PH * PW][kTileC];
  __shared__ uint8_t inside[PH * PW];
  const int C = p.c0 + p.c1;
  const int cblocks = C / kTileC;
  int bid = blockIdx.x;
  const int cb = bid % cblocks; bid /= cblocks;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  const int c0 = cb * kTileC;
  const int oy0 = ty * kTileH, ox0 = tx * kTileW;
  const int tid = threadIdx.x;

  // ---- phase 1: patch -> smem (fp32, after GroupNorm + SiLU); thread = fixed 4 channels, strided pixels
  {
    const int c4 = (tid & 15) * 4;
    const int c = c0 + c4;
    float ga[4] = {1.f, 1.f, 1.f, 1.f}, gs[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.gn_mean != nullptr) {
      const int cpg = C >> 5;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int g = (c + i) / cpg;
        const float mean = __ldg(p.gn_mean + b * 32 + g), rstd = __ldg(p.gn_rstd + b * 32 + g);
        ga[i] = rstd * __ldg(p.gn_gamma + c + i);
        gs[i] = fmaf(-mean, ga[i], __ldg(p.gn_beta + c + i));
      }
    }
    const TIn* src = c < p.c0 ? static_cast<const TIn*>(p.src0) + c : static_cast<const TIn*>(p.src1) + (c - p.c0);
    const int cs = c < p.c0 ? p.c0 : p.c1;
    for (int pp = tid >> 4; pp < PH * PW; pp += 16) {
      const int py = pp / PW, px = pp % PW;
      const int iy = oy0 + py - KS / 2, ix = ox0 + px - KS / 2;
      const bool in = iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) {
        const int sy = p.upsample ? (iy >> 1) : iy, sx = p.upsample ? (ix >> 1) : ix;
        const size_t pix = (static_cast<size_t>(b) * p.hs + sy) * p.ws + sx;
        float u[4];
        if (sizeof(TIn) == 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + pix * cs);
          u[0] = q4.x; u[1] = q4.y; u[2] = q4.z; u[3] = q4.w;
        } else {
          const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(src) + pix * cs);
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
          const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
          u[0] = lo.x; u[1] = lo.y; u[2] = hi.x; u[3] = hi.y;
        }
        if (p.gn_mean != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) u[i] = fmaf(u[i], ga[i], gs[i]);
        }
        if (p.act == 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) u[i] = silu_f(u[i]);
        }
        v = make_float4(u[0], u[1], u[2], u[3]);
      }
      *reinterpret_cast<float4*>(&patch[pp][((c4 >> 2) & 1) * (kTileC / 2) + (c4 >> 3) * 4]) = v;
      if ((tid & 15) == 0) inside[pp] = in ? 1 : 0;
    }
  }
  __syncthreads();

  // ---- phase 2: thread = fixed 8 channels; per tap: cache (delta, 1/delta, zp), sweep 4 pixels
  const int cg = (tid & 7) * 8;
  const int ps = tid >> 3;  // 0..31
  const QuantDev& q = p.q;
  const bool emit_int = q.emit_int != 0;
  for (int tap = 0; tap < KS * KS; ++tap) {
    const int dy = tap / KS, dx = tap % KS;
    const int k0 = tap * C + c0 + cg;
    float d[8], inv[8], z[8];
    if (QMODE == DGQ_Q_KWISE) {
      ldg8(q.delta + k0, d);
      ldg8(q.zp + k0, z);
      if (q.inv != nullptr) {
        ldg8(q.inv + k0, inv);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) inv[i] = rcp_rn_slow(d[i]);
      }
    } else if (QMODE == DGQ_Q_SCALAR) {
      const float dd = __ldg(q.delta), zz = __ldg(q.zp), ii = __frcp_rn(dd);
#pragma unroll
      for (int i = 0; i < 8; ++i) { d[i] = dd; z[i] = zz; inv[i] = ii; }
    }
    float qlo[8], qhi[8];                 // clamp bounds of (code - zp), once per tap
    if (QMODE == DGQ_Q_KWISE && !kCodes) uaq_bounds<8>(z, q.qmax, qlo, qhi);
#pragma unroll
    for (int it = 0; it < (kTileH * kTileW) / 32; ++it) {
      const int pl = ps + it * 32;
      const int oy = oy0 + pl / kTileW, ox = ox0 + pl % kTileW;
      if (oy >= p.ho || ox >= p.wo) continue;
      const int pp = (pl / kTileW + dy) * PW + (pl % kTileW + dx);
      const int m = (b * p.ho + oy) * p.wo + ox;
      float v[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&patch[pp][cg >> 1]);
      const float4 a1 = *reinterpret_cast<const float4*>(&patch[pp][(kTileC / 2) + (cg >> 1)]);
      v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
      const bool quantize = QMODE != DGQ_Q_NONE && (inside[pp] || p.pad_quantized);
      float zrow = 0.f;                   // zero point of this row, for the u8-code output
      if (kCodes) {
        if (QMODE == DGQ_Q_ROWWISE) {
          const int j = m % q.period;
          const float dd = __ldg(q.delta + j), zz = __ldg(q.zp + j), ii = __frcp_rn(dd);
#pragma unroll
          for (int i = 0; i < 8; ++i) { d[i] = dd; z[i] = zz; inv[i] = ii; }
        }
        uint32_t lo = 0, hi = 0;
        if (quantize) {
          float cd[8];
          uaq_codes_rcp<8>(v, d, inv, z, q.qmax, cd);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = emit_int ? __fsub_rn(cd[i], z[i]) : uaq_dequant(cd[i], d[i], z[i]);
            if (i < 4) lo |= static_cast<uint32_t>(cd[i]) << (8 * i);
            else hi |= static_cast<uint32_t>(cd[i]) << (8 * (i - 4));
          }
        }
        zrow = z[0];
        *reinterpret_cast<uint2*>(p.codes + static_cast<size_t>(m) * (KS * KS * C) + k0) = make_uint2(lo, hi);
      } else if (quantize) {
        if (QMODE == DGQ_Q_KWISE) {
          uaq_lean_lh<false, 8>(v, d, inv, qlo, qhi);
        } else {
          float dd = d[0], zz = z[0], ii = inv[0];
          if (QMODE == DGQ_Q_ROWWISE) {
            const int j = m % q.period;
            dd = __ldg(q.delta + j); zz = __ldg(q.zp + j);
            ii = q.inv != nullptr ? __ldg(q.inv + j) : rcp_rn_slow(dd);
          }
          zrow = zz;
          if (emit_int) uaq_lean1_lh<true, 8>(v, dd, ii, -zz, __fsub_rn(q.qmax, zz));
          else uaq_lean1_lh<false, 8>(v, dd, ii, -zz, __fsub_rn(q.qmax, zz));
        }
      } else if (QMODE == DGQ_Q_SCALAR) {
        zrow = z[0];                      // padding tap left at exact 0: the u8 operand stores the code zp
      } else if (QMODE == DGQ_Q_ROWWISE && q.emit_int == 2) {
        zrow = __ldg(q.zp + m % q.period);
      }
      store8(p.out, static_cast<size_t>(m) * p.ldo + k0, v, q.emit_int, zrow);
    }
  }
}

template <typename TIn, int KS>
static void launch_conv_producer(const ProducerDev& p, cudaStream_t s) {
  const int tiles_x = (p.wo + kTileW - 1) / kTileW, tiles_y = (p.ho + kTileH - 1) / kTileH;
  const int grid = p.batch * tiles_y * tiles_x * ((p.c0 + p.c1) / kTileC);
  if (p.codes != nullptr) {   // verification path: also emits the integer codes
    switch (p.q.mode) {
      case DGQ_Q_KWISE: conv_producer_kernel<TIn, KS, DGQ_Q_KWISE, true><<<grid, 256, 0, s>>>(p, tiles_x, tiles_y); break;
      case DGQ_Q_ROWWISE: conv_producer_kernel<TIn, KS, DGQ_Q_ROWWISE, true><<<grid, 256, 0, s>>>(p, tiles_x, tiles_y); break;
      case DGQ_Q_SCALAR: conv_producer_kernel<TIn, KS, DGQ_Q_SCALAR, true><<<grid, 256, 0, s>>>(p, tiles_x, tiles_y); break;
      default: conv_producer_kernel<TIn, KS, DGQ_Q_NONE, true><<<grid, 256, 0, s>>>(p, tiles_x, tiles_y); break;
    }
    return;
  }
  switch (p.q.mode) {
    case DGQ_Q_KWISE: conv_producer_kernel<TIn, KS, DGQ_Q_KWISE, false><<<grid, 256, 0, s>>>(p, tiles_x, tiles_y); break;
    case DGQ_Q_ROWWISE: conv_producer_kernel<TIn, KS, DGQ_Q_ROWWISE, false><<<grid, 256, 0, s>>>(p, tiles_x, tiles_y); break;
    case DGQ_Q_SCALAR: conv_producer_kernel<TIn, KS, DGQ_Q_SCALAR, false><<<grid, 256, 0, s>>>(p, tiles_x, tiles_y); break;
    default: conv_producer_kernel<TIn, KS, DGQ_Q_NONE, false><<<grid, 256, 0, s>>>(p, tiles_x, tiles_y); break;
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm statistics, stage 1: CTA (b, chunk) sums rows [chunk*rows_per, ...) per channel, folds
// channels into the 32 groups, writes partial (sum, sumsq) to scratch[b][chunk][32][2].
constexpr int kGnMaxC = 2560;
template <typename TIn>
__global__ void __launch_bounds__(256) gn_partial_kernel(const TIn* __restrict__ src0,
                                                         const TIn* __restrict__ src1, int c0, int c1, int hw,
                                                         int rows_per, int chunks, float* __restrict__ scratch) {
  const int C = c0 + c1;
  const int cpg = C >> 5;
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int r0 = chunk * rows_per;
  const int r1 = min(hw, r0 + rows_per);
  // per-(row-lane, channel) partial sums; reduced in a fixed order below (deterministic)
  __shared__ float s_sum[kGnMaxC], s_sq[kGnMaxC];
  const int cvec = C >> 3;
  const int tid = threadIdx.x;
  int cv0, sub, nsub, cstep;
  if (cvec >= static_cast<int>(blockDim.x)) { cv0 = tid; sub = 0; nsub = 1; cstep = blockDim.x; }
  else { nsub = blockDim.x / cvec; cv0 = tid % cvec; sub = tid / cvec; cstep = cvec; if (sub >= nsub) cv0 = cvec; }
  for (int cv = cv0; cv < cvec; cv += cstep) {
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
    const int c = cv << 3;
    const TIn* base = c < c0 ? src0 + c : src1 + (c - c0);
    const size_t cs = c < c0 ? c0 : c1;
    int r = r0 + sub;
    for (; r + 3 * nsub < r1; r += 4 * nsub) {      // four independent row loads in flight per thread
      float v[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) load8(base + (static_cast<size_t>(b) * hw + r + k * nsub) * cs, v[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += v[k][i]; q[i] += v[k][i] * v[k][i]; }
      }
    }
    for (; r < r1; r += nsub) {
      float v[8];
      load8(base + (static_cast<size_t>(b) * hw + r) * cs, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += v[i]; q[i] += v[i] * v[i]; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { s_sum[sub * C + c + i] = s[i]; s_sq[sub * C + c + i] = q[i]; }
  }
  __syncthreads();
  if (tid < 32) {
    float s = 0.f, q = 0.f;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c)
      for (int u = 0; u < nsub; ++u) { s += s_sum[u * C + c]; q += s_sq[u * C + c]; }
    float* o = scratch + ((static_cast<size_t>(b) * chunks + chunk) * 32 + tid) * 2;
    o[0] = s;
    o[1] = q;
  }
}
__global__ void gn_final_kernel(const float* __restrict__ scratch, int chunks, double count, float eps,
                                float* __restrict__ mean, float* __restrict__ rstd, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (b, g)
  if (i >= total) return;
  const int b = i >> 5, g = i & 31;
  double s = 0.0, q = 0.0;
  for (int ch = 0; ch < chunks; ++ch) {
    const float* o = scratch + ((static_cast<size_t>(b) * chunks + ch) * 32 + g) * 2;
    s += o[0];
    q += o[1];
  }
  const double mu = s / count;
  double var = q / count - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[i] = static_cast<float>(mu);
  rstd[i] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// ------------------------------------------------------------------------------------------
// LayerNorm / plain rows + up to 3 quantizers.  One warp per row; row cached in registers.
struct RowQuantDev {
  QuantDev q[3];
  __half* out[3];
  uint8_t* codes[3];
  int n_out;
};

// kMaxVecPerLane * 256 = widest supported row (5: C <= 1280, the LayerNorm widths; 12: C <= 3072)
template <typename TIn, bool kNorm, int kMaxVecPerLane>
__global__ void __launch_bounds__(256) row_quant_kernel(const TIn* __restrict__ x, int m, int c,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        const RowQuantDev rq) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= m) return;
  const int cvec = c >> 3;
  float v[kMaxVecPerLane][8];
  const TIn* row = x + static_cast<size_t>(warp) * c;
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVecPerLane; ++j) {
    const int cv = lane + j * 32;
    if (cv < cvec) {
      load8(row + (cv << 3), v[j]);
      if (kNorm) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += v[j][i];
      }
    }
  }
  if (kNorm) {
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / static_cast<float>(c);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxVecPerLane; ++j) {
      if (lane + j * 32 < cvec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; sq += d * d; }
      }
    }
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.0f / sqrtf(sq / static_cast<float>(c) + eps);
#pragma unroll
    for (int j = 0; j < kMaxVecPerLane; ++j) {
      const int cv = lane + j * 32;
      if (cv < cvec) {
        float ga[8], be[8];
        ldg8(gamma + (cv << 3), ga);
        ldg8(beta + (cv << 3), be);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j][i] = (v[j][i] - mean) * rstd * ga[i] + be[i];
      }
    }
  }
#pragma unroll 1
  for (int o = 0; o < rq.n_out; ++o) {
    const QuantDev q = rq.q[o];
    __half* orow = rq.out[o] + static_cast<size_t>(warp) * c;
    const float zrow = q.emit_int == 2 ? quant_zp(q, warp) : 0.0f;
    if (rq.codes[o] == nullptr) {
#pragma unroll
      for (int j = 0; j < kMaxVecPerLane; ++j) {
        const int cv = lane + j * 32;
        if (cv < cvec) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = v[j][i];
          quant8_lean(q, t, cv << 3, warp);
          store8(rq.out[o], static_cast<size_t>(warp) * c + (cv << 3), t, q.emit_int, zrow);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < kMaxVecPerLane; ++j) {
        const int cv = lane + j * 32;
        if (cv < cvec) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = v[j][i];
          quant8(q, t, cv << 3, warp, rq.codes[o] + static_cast<size_t>(warp) * c + (cv << 3));
          *reinterpret_cast<uint4*>(orow + (cv << 3)) = pack8(t);
        }
      }
    }
  }
}

// rows wider than the register-cached kernel supports (ff.net.2 inputs, c = 5120, when they do not come from the
// fused GEGLU epilogue): quantizer only, streamed -- one warp per row, nothing cached
template <typename TIn>
__global__ void __launch_bounds__(256) row_quant_wide_kernel(const TIn* __restrict__ x, int m, int c, const RowQuantDev rq) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= m) return;
  const TIn* row = x + static_cast<size_t>(warp) * c;
  for (int o = 0; o < rq.n_out; ++o) {
    const QuantDev q = rq.q[o];
    for (int cv = lane; cv < (c >> 3); cv += 32) {
      float t[8];
      load8(row + (cv << 3), t);
      if (rq.codes[o] == nullptr) quant8_lean(q, t, cv << 3, warp);
      else quant8(q, t, cv << 3, warp, rq.codes[o] + static_cast<size_t>(warp) * c + (cv << 3));
      store8(rq.out[o], static_cast<size_t>(warp) * c + (cv << 3), t, q.emit_int, q.emit_int == 2 ? quant_zp(q, warp) : 0.0f);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Tiled LayerNorm / row quantizer for the hot widths (c = 320 / 640 / 1280): a CTA owns kLnRows rows.
//   phase 1  one warp per row (two rows per warp): load, mean / rstd, normalise -> fp32 tile in smem
//   phase 2  thread = (8 columns, a slice of the rows): the K-wise (delta, 1/delta, zp) of its 8 columns are
//            loaded ONCE per quantizer and applied to every row of its slice
// The row-per-warp kernel above re-reads 3 tables x 3 quantizers + gamma / beta through L1 for every row
// (11 x the activation bytes: ncu shows it L1-bound at 72 % with DRAM at 22 %).  Same arithmetic, so the
// results are bit-identical to it.
constexpr int kLnThreads = 320, kLnRows = 20;

template <typename TIn, bool kNorm>
__global__ void __launch_bounds__(kLnThreads, 2) ln_tile_kernel(const TIn* __restrict__ x, int m, int c,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps,
                                                                const RowQuantDev rq) {
  extern __shared__ __align__(16) float ln_tile[];   // [kLnRows][2 planes][cvec] float4: plane p = columns 4p .. 4p+3 of each 8
  constexpr int kMaxVec = 5;                          // c <= 1280
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kLnRows;
  const int cvec = c >> 3;
  float4* tile4 = reinterpret_cast<float4*>(ln_tile);

  for (int rr = warp; rr < kLnRows; rr += kLnThreads / 32) {
    const int row_i = row0 + rr;
    if (row_i >= m) break;
    float v[kMaxVec][8];
    const TIn* row = x + static_cast<size_t>(row_i) * c;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int cv = lane + j * 32;
      if (cv < cvec) {
        load8(row + (cv << 3), v[j]);
        if (kNorm) {
#pragma unroll
          for (int i = 0; i < 8; ++i) sum += v[j][i];
        }
      }
    }
    if (kNorm) {
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / static_cast<float>(c);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) {
        if (lane + j * 32 < cvec) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; sq += d * d; }
        }
      }
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = 1.0f / sqrtf(sq / static_cast<float>(c) + eps);
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) {
        const int cv = lane + j * 32;
        if (cv < cvec) {
          float ga[8], be[8];
          ldg8(gamma + (cv << 3), ga);
          ldg8(beta + (cv << 3), be);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j][i] = (v[j][i] - mean) * rstd * ga[i] + be[i];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int cv = lane + j * 32;
      if (cv < cvec) {
        tile4[(rr * 2 + 0) * cvec + cv] = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
        tile4[(rr * 2 + 1) * cvec + cv] = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
      }
    }
  }
  __syncthreads();

  const int groups = kLnThreads / cvec;               // cvec divides kLnThreads (checked by the launcher)
  const int cv = threadIdx.x % cvec, grp = threadIdx.x / cvec;
  const int rpg = (kLnRows + groups - 1) / groups;
  const int r_begin = grp * rpg;
  int r_end = r_begin + rpg;
  if (r_end > kLnRows) r_end = kLnRows;
  if (r_end > m - row0) r_end = m - row0;
  const int k0 = cv << 3;
#pragma unroll 1
  for (int o = 0; o < rq.n_out; ++o) {
    const QuantDev q = rq.q[o];
    __half* obase = rq.out[o] + static_cast<size_t>(row0) * c + k0;
    if (q.mode == DGQ_Q_KWISE) {
      float d[8], z[8], inv[8];
      ldg8(q.delta + k0, d);
      ldg8(q.zp + k0, z);
      if (q.inv != nullptr) {
        ldg8(q.inv + k0, inv);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) inv[i] = rcp_rn_slow(d[i]);
      }
      float qlo[8], qhi[8];
      uaq_bounds<8>(z, q.qmax, qlo, qhi);
      for (int r = r_begin; r < r_end; ++r) {
        const float4 a0 = tile4[(r * 2 + 0) * cvec + cv], a1 = tile4[(r * 2 + 1) * cvec + cv];
        float t[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        uaq_lean_lh<false, 8>(t, d, inv, qlo, qhi);
        *reinterpret_cast<uint4*>(obase + static_cast<size_t>(r) * c) = pack8(t);
      }
    } else if (q.mode == DGQ_Q_SCALAR) {
      // per-tensor scale (every layer of the g = 1 configs): the triple and the clamp bounds once per quantizer, not
      // once per row (the generic path below re-read them through L1 for each row, with a division when q.inv is absent)
      const float dd = __ldg(q.delta), zz = __ldg(q.zp);
      const float ii = q.inv != nullptr ? __ldg(q.inv) : rcp_rn_slow(dd);
      const float lo = -zz, hi = __fsub_rn(q.qmax, zz);
      for (int r = r_begin; r < r_end; ++r) {
        const float4 a0 = tile4[(r * 2 + 0) * cvec + cv], a1 = tile4[(r * 2 + 1) * cvec + cv];
        float t[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (q.emit_int) uaq_lean1_lh<true, 8>(t, dd, ii, lo, hi);
        else uaq_lean1_lh<false, 8>(t, dd, ii, lo, hi);
        store8(rq.out[o], static_cast<size_t>(row0 + r) * c + k0, t, q.emit_int, zz);
      }
    } else {
      for (int r = r_begin; r < r_end; ++r) {
        const float4 a0 = tile4[(r * 2 + 0) * cvec + cv], a1 = tile4[(r * 2 + 1) * cvec + cv];
        float t[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        quant8_lean(q, t, k0, row0 + r);
        store8(rq.out[o], static_cast<size_t>(row0 + r) * c + k0, t, q.emit_int,
               q.emit_int == 2 ? quant_zp(q, row0 + r) : 0.0f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) geglu_quant_kernel(const TIn* __restrict__ x, int m, int f,
                                                          const QuantDev q, __half* __restrict__ out) {
  const int fvec = f >> 3;
  const int64_t total = static_cast<int64_t>(m) * fvec;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(idx / fvec);
    const int k0 = static_cast<int>(idx % fvec) << 3;
    float a[8], g[8];
    load8(x + static_cast<size_t>(row) * 2 * f + k0, a);
    load8(x + static_cast<size_t>(row) * 2 * f + f + k0, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * gelu_erf_f(g[i]);
    quant8_lean(q, a, k0, row);
    store8(out, static_cast<size_t>(row) * f + k0, a, q.emit_int, q.emit_int == 2 ? quant_zp(q, row) : 0.0f);
  }
}

// ------------------------------------------------------------------------------------------
// x [b*t, ldx] (head h = columns h*d..) -> out [b, heads, t, dp] (transpose == 0)
//                                        or out [b, heads, dp, tp] (transpose == 1)
template <typename TIn>
__global__ void __launch_bounds__(256) qkv_pack_kernel(const TIn* __restrict__ x, int ldx, int b, int t,
                                                       int heads, int d, int dp, int tp, int transpose,
                                                       int skip_first, const QuantDev q, const float* __restrict__ kfold,
                                                       int k_split, __half* __restrict__ out) {
  if (!transpose) {
    const int dvec = dp >> 3;
    const int ldk = k_split ? 2 * dp : dp;
    const int64_t total = static_cast<int64_t>(b) * heads * t * dvec;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int d0 = static_cast<int>(idx % dvec) << 3;
      int64_t r = idx / dvec;
      const int tt = static_cast<int>(r % t); r /= t;
      const int hh = static_cast<int>(r % heads);
      const int bb = static_cast<int>(r / heads);
      float v[8];
      if (d0 < d) {
        load8(x + (static_cast<size_t>(bb) * t + tt) * ldx + hh * d + d0, v);
        if (!(skip_first && tt == 0)) {
          // KWISE index = d, ROWWISE index = token (minus the bypassed start token); emit_int: the bare integer
          // code - zp in every mode (the scale is applied by the attention kernel / folded into K)
          QuantDev qq = q;
          if (qq.mode == DGQ_Q_ROWWISE) qq.period = 1 << 30;
          quant8(qq, v, d0, tt - skip_first, nullptr);
        }
        if (kfold != nullptr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= __ldg(kfold + d0 + i);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      __half* dst = out + ((static_cast<size_t>(bb) * heads + hh) * t + tt) * ldk + d0;
      const uint4 hi = pack8(v);
      *reinterpret_cast<uint4*>(dst) = hi;
      if (k_split) {
        Half8 h;
        h.u = hi;
        float lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) lo[i] = v[i] - __half2float(h.h[i]);
        *reinterpret_cast<uint4*>(dst + dp) = pack8(lo);
      }
    }
  } else {
    const int tvec = tp >> 3;
    const int64_t total = static_cast<int64_t>(b) * heads * dp * tvec;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      // dd fastest so that a warp reads consecutive channels of the same token rows
      const int dd = static_cast<int>(idx % dp);
      int64_t r = idx / dp;
      const int t0 = static_cast<int>(r % tvec) << 3; r /= tvec;
      const int hh = static_cast<int>(r % heads);
      const int bb = static_cast<int>(r / heads);
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int tt = t0 + i;
        float val = 0.f;
        if (dd < d && tt < t) {
          val = to_f(x[(static_cast<size_t>(bb) * t + tt) * ldx + hh * d + dd]);
          if (q.mode != DGQ_Q_NONE && !(skip_first && tt == 0)) {
            const int j = q.mode == DGQ_Q_KWISE ? dd : (q.mode == DGQ_Q_ROWWISE ? tt - skip_first : 0);
            const float dl = __ldg(q.delta + j), z = __ldg(q.zp + j);
            val = uaq_dequant(uaq_code(val, dl, z, q.qmax), dl, z);
          }
        }
        v[i] = val;
      }
      *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(bb) * heads + hh) * dp + dd) * tp + t0) = pack8(v);
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int n, int dim, __half* __restrict__ o16,
                                          float* __restrict__ o32, int ldo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim >> 1;
  if (i >= n * half) return;
  const int r = i / half, j = i % half;
  // exponent = -ln(10000) * j / half   (diffusers_rewrite/sd.py:27-31)
  const float e = expf(-9.210340371976184f * static_cast<float>(j) / static_cast<float>(half));
  const float a = t[r] * e;
  const float cs = cosf(a), sn = sinf(a);
  if (o16 != nullptr) {
    o16[static_cast<size_t>(r) * ldo + j] = __float2half_rn(cs);
    o16[static_cast<size_t>(r) * ldo + half + j] = __float2half_rn(sn);
  }
  if (o32 != nullptr) {
    o32[static_cast<size_t>(r) * ldo + j] = cs;
    o32[static_cast<size_t>(r) * ldo + half + j] = sn;
  }
}

template <typename TOut>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int b, int c, int hw, int c_pad,
                                    TOut* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(b) * hw * c_pad;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c_pad);
    const int64_t r = i / c_pad;
    const int p = static_cast<int>(r % hw);
    const int bb = static_cast<int>(r / hw);
    from_f(cc < c ? x[(static_cast<size_t>(bb) * c + cc) * hw + p] : 0.f, out + i);
  }
}
template <typename TIn>
__global__ void nhwc_to_nchw_kernel(const TIn* __restrict__ x, int b, int c, int hw, int ldx,
                                    float* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(b) * c * hw;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i % hw);
    const int64_t r = i / hw;
    const int cc = static_cast<int>(r % c);
    const int bb = static_cast<int>(r / c);
    out[i] = to_f(x[(static_cast<size_t>(bb) * hw + p) * ldx + cc]);
  }
}
template <typename T>
__global__ void silu_kernel(const T* __restrict__ x, int64_t n, T* __restrict__ out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = to_f(x[i]);
    from_f(v / (1.0f + expf(-v)), out + i);
  }
}
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n, T* __restrict__ out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    from_f(to_f(a[i]) + to_f(b[i]), out + i);
}

static int grid_for(int64_t work, int block, int max_blocks) {
  int64_t g = (work + block - 1) / block;
  if (g < 1) g = 1;
  return static_cast<int>(g < max_blocks ? g : max_blocks);
}
static bool quant_ok(const dgq_quant_t& q) {
  if (q.mode == DGQ_Q_NONE) return q.emit_int != 2;   // u8 codes need a quantizer
  if (q.mode < 0 || q.mode > DGQ_Q_ROWWISE) return false;
  if (q.delta == nullptr || q.zp == nullptr) return false;
  if (q.mode == DGQ_Q_ROWWISE && q.period <= 0) return false;
  if (q.emit_int && q.mode == DGQ_Q_KWISE) return false;
  if (q.emit_int < 0 || q.emit_int > 2) return false;
  return true;
}

}  // namespace dgq

extern "C" int dgq_act_producer(const dgq_producer_t* a, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(a != nullptr && a->src0 != nullptr && a->out != nullptr);
  DGQ_CHECK_ARG(a->c0 > 0 && a->c0 % 8 == 0 && a->c1 >= 0 && a->c1 % 8 == 0);
  DGQ_CHECK_ARG(a->c1 == 0 || a->src1 != nullptr);
  DGQ_CHECK_ARG(a->batch > 0 && a->h > 0 && a->w > 0);
  DGQ_CHECK_ARG((a->ksize == 1 && a->pad == 0) || (a->ksize == 3 && a->pad == 1));
  DGQ_CHECK_ARG(a->stride == 1 || a->stride == 2);
  DGQ_CHECK_ARG(!a->upsample || (a->h % 2 == 0 && a->w % 2 == 0));
  DGQ_CHECK_ARG(quant_ok(a->q));
  const int C = a->c0 + a->c1;
  const int K = a->ksize * a->ksize * C;
  DGQ_CHECK_ARG(a->ldo >= K && a->ldo % 8 == 0);
  DGQ_CHECK_ARG(a->gn_mean == nullptr ||
                (a->gn_rstd != nullptr && a->gn_gamma != nullptr && a->gn_beta != nullptr && C % 32 == 0));
  ProducerDev p;
  p.src0 = a->src0; p.src1 = a->src1; p.c0 = a->c0; p.c1 = a->c1;
  p.batch = a->batch; p.h = a->h; p.w = a->w;
  p.hs = a->upsample ? a->h / 2 : a->h; p.ws = a->upsample ? a->w / 2 : a->w;
  p.ho = (a->h + 2 * a->pad - a->ksize) / a->stride + 1;
  p.wo = (a->w + 2 * a->pad - a->ksize) / a->stride + 1;
  p.upsample = a->upsample; p.ksize = a->ksize; p.stride = a->stride; p.pad = a->pad;
  p.gn_mean = a->gn_mean; p.gn_rstd = a->gn_rstd; p.gn_gamma = a->gn_gamma; p.gn_beta = a->gn_beta;
  p.act = a->act; p.q = to_dev(a->q); p.pad_quantized = a->pad_quantized;
  p.out = static_cast<__half*>(a->out); p.ldo = a->ldo; p.codes = a->codes;
  const int64_t total = static_cast<int64_t>(p.batch) * p.ho * p.wo * (p.ldo / 8);
  const int grid = grid_for(total, 256, kNumSMs * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // hot shapes: tiled kernel (each input element normalised once, taps emitted from smem)
  if (a->stride == 1 && C % kTileC == 0 && a->c0 % 4 == 0 && p.ldo == K &&
      static_cast<int64_t>(p.batch) * p.ho * p.wo * 9 * C < (int64_t(1) << 40)) {
    if (a->ksize == 3) {
      if (a->src_is_f32) launch_conv_producer<float, 3>(p, s); else launch_conv_producer<__half, 3>(p, s);
    } else {
      if (a->src_is_f32) launch_conv_producer<float, 1>(p, s); else launch_conv_producer<__half, 1>(p, s);
    }
    DGQ_RETURN_LAST_ERROR();
  }
  if (a->src_is_f32) act_producer_kernel<float><<<grid, 256, 0, s>>>(p);
  else act_producer_kernel<__half><<<grid, 256, 0, s>>>(p);
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_gn_stats(const void* src0, const void* src1, int src_is_f32, int c0, int c1, int batch, int hw,
                            float eps, float* mean, float* rstd, float* scratch, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(src0 != nullptr && mean != nullptr && rstd != nullptr && scratch != nullptr);
  DGQ_CHECK_ARG(c0 > 0 && c0 % 8 == 0 && c1 >= 0 && c1 % 8 == 0 && (c0 + c1) % 32 == 0);
  DGQ_CHECK_ARG(c1 == 0 || src1 != nullptr);
  DGQ_CHECK_ARG(batch > 0 && hw > 0 && (c0 + c1) <= kGnMaxC);
  int chunks = (hw + 15) / 16;           // >= 4 CTAs per SM at the SDXL sizes (scratch holds 64 chunks per sample)
  if (chunks > 64) chunks = 64;
  const int rows_per = (hw + chunks - 1) / chunks;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (src_is_f32)
    gn_partial_kernel<float><<<batch * chunks, 256, 0, s>>>(static_cast<const float*>(src0),
                                                            static_cast<const float*>(src1), c0, c1, hw, rows_per,
                                                            chunks, scratch);
  else
    gn_partial_kernel<__half><<<batch * chunks, 256, 0, s>>>(static_cast<const __half*>(src0),
                                                             static_cast<const __half*>(src1), c0, c1, hw, rows_per,
                                                             chunks, scratch);
  const double count = static_cast<double>(hw) * ((c0 + c1) / 32);
  gn_final_kernel<<<(batch * 32 + 127) / 128, 128, 0, s>>>(scratch, chunks, count, eps, mean, rstd, batch * 32);
  DGQ_RETURN_LAST_ERROR();
}

static int launch_row_quant(const void* x, int src_is_f32, bool norm, int m, int c, const float* gamma,
                            const float* beta, float eps, int n_out, const dgq_quant_t* q, void* const* out,
                            uint8_t* const* codes, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(x != nullptr && m > 0 && c > 0 && c % 8 == 0 && (c <= 12 * 256 || !norm));
  DGQ_CHECK_ARG(n_out >= 1 && n_out <= 3 && q != nullptr && out != nullptr);
  RowQuantDev rq;
  rq.n_out = n_out;
  for (int i = 0; i < 3; ++i) {
    rq.out[i] = nullptr; rq.codes[i] = nullptr;
    rq.q[i] = QuantDev{nullptr, nullptr, nullptr, DGQ_Q_NONE, 1, 0.f, 0};
  }
  for (int i = 0; i < n_out; ++i) {
    DGQ_CHECK_ARG(out[i] != nullptr && quant_ok(q[i]));
    rq.q[i] = to_dev(q[i]);
    rq.out[i] = static_cast<__half*>(out[i]);
    rq.codes[i] = codes != nullptr ? codes[i] : nullptr;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cvec = c / 8;
  if (codes == nullptr && c <= 1280 && kLnThreads % cvec == 0 && m >= 4 * kLnRows) {
    // hot widths: tiled kernel (quantizer tables read once per CTA instead of once per row)
    const int tgrid = (m + kLnRows - 1) / kLnRows;
    const size_t smem = static_cast<size_t>(kLnRows) * c * sizeof(float);
    static PerDeviceOnce attr;
    int dev;
    if (!attr.done(&dev)) {
      const int kMaxSmem = kLnRows * 1280 * 4;
      cudaError_t e = cudaFuncSetAttribute(ln_tile_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(ln_tile_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(ln_tile_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(ln_tile_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
      if (e != cudaSuccess) return static_cast<int>(e);
      attr.mark(dev);
    }
    if (norm) {
      DGQ_CHECK_ARG(gamma != nullptr && beta != nullptr);
      if (src_is_f32) ln_tile_kernel<float, true><<<tgrid, kLnThreads, smem, s>>>(static_cast<const float*>(x), m, c, gamma, beta, eps, rq);
      else ln_tile_kernel<__half, true><<<tgrid, kLnThreads, smem, s>>>(static_cast<const __half*>(x), m, c, gamma, beta, eps, rq);
    } else {
      if (src_is_f32) ln_tile_kernel<float, false><<<tgrid, kLnThreads, smem, s>>>(static_cast<const float*>(x), m, c, nullptr, nullptr, 0.f, rq);
      else ln_tile_kernel<__half, false><<<tgrid, kLnThreads, smem, s>>>(static_cast<const __half*>(x), m, c, nullptr, nullptr, 0.f, rq);
    }
    DGQ_RETURN_LAST_ERROR();
  }
  const int grid = (m + 7) / 8;  // 8 warps (rows) per CTA
  if (c > 12 * 256) {
    if (src_is_f32) row_quant_wide_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), m, c, rq);
    else row_quant_wide_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(x), m, c, rq);
    DGQ_RETURN_LAST_ERROR();
  }
  const bool narrow = c <= 5 * 256;
  if (norm) {
    DGQ_CHECK_ARG(gamma != nullptr && beta != nullptr);
    if (src_is_f32) {
      if (narrow) row_quant_kernel<float, true, 5><<<grid, 256, 0, s>>>(static_cast<const float*>(x), m, c, gamma, beta, eps, rq);
      else row_quant_kernel<float, true, 12><<<grid, 256, 0, s>>>(static_cast<const float*>(x), m, c, gamma, beta, eps, rq);
    } else {
      if (narrow) row_quant_kernel<__half, true, 5><<<grid, 256, 0, s>>>(static_cast<const __half*>(x), m, c, gamma, beta, eps, rq);
      else row_quant_kernel<__half, true, 12><<<grid, 256, 0, s>>>(static_cast<const __half*>(x), m, c, gamma, beta, eps, rq);
    }
  } else if (src_is_f32) {
    if (narrow) row_quant_kernel<float, false, 5><<<grid, 256, 0, s>>>(static_cast<const float*>(x), m, c, nullptr, nullptr, 0.f, rq);
    else row_quant_kernel<float, false, 12><<<grid, 256, 0, s>>>(static_cast<const float*>(x), m, c, nullptr, nullptr, 0.f, rq);
  } else {
    if (narrow) row_quant_kernel<__half, false, 5><<<grid, 256, 0, s>>>(static_cast<const __half*>(x), m, c, nullptr, nullptr, 0.f, rq);
    else row_quant_kernel<__half, false, 12><<<grid, 256, 0, s>>>(static_cast<const __half*>(x), m, c, nullptr, nullptr, 0.f, rq);
  }
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_ln_quant(const void* x, int src_is_f32, int m, int c, const float* gamma, const float* beta,
                            float eps, int n_out, const dgq_quant_t* host_q, void* const* host_out, void* stream) {
  return launch_row_quant(x, src_is_f32, true, m, c, gamma, beta, eps, n_out, host_q, host_out, nullptr, stream);
}
extern "C" int dgq_row_quant(const void* x, int src_is_f32, int m, int c, int n_out, const dgq_quant_t* host_q,
                             void* const* host_out, uint8_t* const* host_codes, void* stream) {
  return launch_row_quant(x, src_is_f32, false, m, c, nullptr, nullptr, 0.f, n_out, host_q, host_out, host_codes,
                          stream);
}

extern "C" int dgq_geglu_quant(const void* x, int src_is_f32, int m, int f, dgq_quant_t q, void* out, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(x != nullptr && out != nullptr && m > 0 && f > 0 && f % 8 == 0 && quant_ok(q));
  const int64_t total = static_cast<int64_t>(m) * (f / 8);
  const int grid = grid_for(total, 256, kNumSMs * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (src_is_f32)
    geglu_quant_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), m, f, to_dev(q), static_cast<__half*>(out));
  else
    geglu_quant_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(x), m, f, to_dev(q), static_cast<__half*>(out));
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_qkv_pack(const void* x, int src_is_f32, int ldx, int b, int t, int heads, int d, int dp, int tp,
                            int transpose, int skip_first, dgq_quant_t q, const float* kfold, int k_split, void* out,
                            void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(x != nullptr && out != nullptr && b > 0 && t > 0 && heads > 0 && d > 0);
  DGQ_CHECK_ARG(d % 8 == 0 && dp >= d && dp % 8 == 0 && ldx % 8 == 0 && q.emit_int >= 0 && q.emit_int <= 1);
  DGQ_CHECK_ARG(q.mode >= DGQ_Q_NONE && q.mode <= DGQ_Q_ROWWISE && (q.mode == DGQ_Q_NONE || (q.delta != nullptr && q.zp != nullptr)));
  DGQ_CHECK_ARG(!(transpose && (k_split || kfold != nullptr || q.emit_int)));
  DGQ_CHECK_ARG(!transpose || (tp >= t && tp % 8 == 0));
  const int64_t total = transpose ? static_cast<int64_t>(b) * heads * dp * (tp / 8)
                                  : static_cast<int64_t>(b) * heads * t * (dp / 8);
  const int grid = grid_for(total, 256, kNumSMs * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (src_is_f32)
    qkv_pack_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), ldx, b, t, heads, d, dp, tp, transpose,
                                                skip_first, to_dev(q), kfold, k_split, static_cast<__half*>(out));
  else
    qkv_pack_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(x), ldx, b, t, heads, d, dp, tp,
                                                 transpose, skip_first, to_dev(q), kfold, k_split, static_cast<__half*>(out));
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_timestep_embedding(const float* t, int n, int dim, void* out_f16, float* out_f32, int ldo,
                                      void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(t != nullptr && n > 0 && dim > 0 && dim % 2 == 0 && ldo >= dim);
  DGQ_CHECK_ARG(out_f16 != nullptr || out_f32 != nullptr);
  const int total = n * (dim / 2);
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      t, n, dim, static_cast<__half*>(out_f16), out_f32, ldo);
  DGQ_RETURN_LAST_ERROR();
}
extern "C" int dgq_nchw_to_nhwc(const float* x, int b, int c, int hw, int c_pad, void* out, int out_is_f32,
                                void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(x != nullptr && out != nullptr && b > 0 && c > 0 && hw > 0 && c_pad >= c);
  const int64_t total = static_cast<int64_t>(b) * hw * c_pad;
  const int grid = grid_for(total, 256, kNumSMs * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (out_is_f32) nchw_to_nhwc_kernel<float><<<grid, 256, 0, s>>>(x, b, c, hw, c_pad, static_cast<float*>(out));
  else nchw_to_nhwc_kernel<__half><<<grid, 256, 0, s>>>(x, b, c, hw, c_pad, static_cast<__half*>(out));
  DGQ_RETURN_LAST_ERROR();
}
extern "C" int dgq_nhwc_to_nchw(const void* x, int src_is_f32, int b, int c, int hw, int ldx, float* out,
                                void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(x != nullptr && out != nullptr && b > 0 && c > 0 && hw > 0 && ldx >= c);
  const int64_t total = static_cast<int64_t>(b) * hw * c;
  const int grid = grid_for(total, 256, kNumSMs * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (src_is_f32) nhwc_to_nchw_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), b, c, hw, ldx, out);
  else nhwc_to_nchw_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(x), b, c, hw, ldx, out);
  DGQ_RETURN_LAST_ERROR();
}
// ---------------------------------------------------------------------------------------------
// P[r, :] = softmax(scale * S[r, :]) as fp16: the attention map of the VAE decoder's single-head mid-block attention
// (F.scaled_dot_product_attention at diffusers/models/attention_processor.py:1244; head dim 512 does not fit the
// TMEM-resident flash kernel, and the block runs once per image).  One CTA per row; the row is read three times
// (max, sum, write) -- the 2nd and 3rd reads hit L1 / L2.
namespace dgq {
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, int cols, int64_t lds,
                                                            float scale_log2e, __half* __restrict__ out, int64_t ldo) {
  __shared__ float red[8];
  const float4* row = reinterpret_cast<const float4*>(s + static_cast<int64_t>(blockIdx.x) * lds);
  const int n4 = cols >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -3.0e38f;
  for (int i = threadIdx.x; i < n4; i += 256) {
    const float4 v = __ldg(row + i);
    m = fmaxf(fmaxf(fmaxf(m, v.x), fmaxf(v.y, v.z)), v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  const float mb = m * scale_log2e;
  float sum = 0.0f;
  for (int i = threadIdx.x; i < n4; i += 256) {
    const float4 v = __ldg(row + i);
    sum += exp2f(fmaf(v.x, scale_log2e, -mb)) + exp2f(fmaf(v.y, scale_log2e, -mb)) +
           exp2f(fmaf(v.z, scale_log2e, -mb)) + exp2f(fmaf(v.w, scale_log2e, -mb));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<int64_t>(blockIdx.x) * ldo);
  for (int i = threadIdx.x; i < n4; i += 256) {
    const float4 v = __ldg(row + i);
    const __half2 a = __floats2half2_rn(exp2f(fmaf(v.x, scale_log2e, -mb)) * inv, exp2f(fmaf(v.y, scale_log2e, -mb)) * inv);
    const __half2 b = __floats2half2_rn(exp2f(fmaf(v.z, scale_log2e, -mb)) * inv, exp2f(fmaf(v.w, scale_log2e, -mb)) * inv);
    orow[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}
}  // namespace dgq

extern "C" int dgq_softmax_rows(const float* s, int64_t rows, int cols, int64_t lds, float scale, void* out_f16,
                                int64_t ldo, void* stream) {
  DGQ_CHECK_ARG(s != nullptr && out_f16 != nullptr && rows > 0 && rows < (1ll << 31) && cols > 0 && cols % 4 == 0);
  DGQ_CHECK_ARG(lds >= cols && lds % 4 == 0 && ldo >= cols && ldo % 4 == 0);
  DGQ_CHECK_ARG((reinterpret_cast<uintptr_t>(s) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_f16) & 7) == 0);
  dgq::softmax_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      s, cols, lds, scale * 1.4426950408889634f, static_cast<__half*>(out_f16), ldo);
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_silu(const void* x, int is_f32, int64_t n, void* out, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(x != nullptr && out != nullptr && n > 0);
  const int grid = grid_for(n, 256, kNumSMs * 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_f32) silu_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), n, static_cast<float*>(out));
  else silu_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(x), n, static_cast<__half*>(out));
  DGQ_RETURN_LAST_ERROR();
}
extern "C" int dgq_add(const void* a, const void* b, int is_f32, int64_t n, void* out, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(a != nullptr && b != nullptr && out != nullptr && n > 0);
  const int grid = grid_for(n, 256, kNumSMs * 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_f32) add_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(a), static_cast<const float*>(b), n, static_cast<float*>(out));
  else add_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(a), static_cast<const __half*>(b), n, static_cast<__half*>(out));
  DGQ_RETURN_LAST_ERROR();
}
