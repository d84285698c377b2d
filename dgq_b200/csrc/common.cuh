// Shared host/device helpers for the dgq_b200 kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dgq_b200.h"

#define DGQ_CHECK_ARG(cond)                    \
  do {                                         \
    if (!(cond)) return DGQ_ERR_INVALID_VALUE; \
  } while (0)

#define DGQ_RETURN_LAST_ERROR()                    \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    return e__ == cudaSuccess ? 0 : static_cast<int>(e__); \
  } while (0)

namespace dgq {

constexpr int kNumSMs = 148;

// cudaFuncSetAttribute (the > 48 KB dynamic shared memory opt-in) is per DEVICE: one flag word per call site, one bit
// per device ordinal.  Racing host threads may both set the attribute (idempotent) before the bit is published.
struct PerDeviceOnce {
  unsigned long long mask = 0;
  bool done(int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    return (__atomic_load_n(&mask, __ATOMIC_ACQUIRE) >> (dev & 63)) & 1ull;
  }
  void mark(int dev) { __atomic_fetch_or(&mask, 1ull << (dev & 63), __ATOMIC_RELEASE); }
};

// quantize one value exactly as UniformAffineQuantizer does (quant/quant_layer.py:295-299):
// IEEE division, round-half-to-even, clamp to [0, level-1].  Returns the integer code as float.
__device__ __forceinline__ float uaq_code(float x, float delta, float zp, float qmax) {
  float q = rintf(__fdiv_rn(x, delta)) + zp;
  return fminf(fmaxf(q, 0.0f), qmax);
}
__device__ __forceinline__ float uaq_dequant(float code, float delta, float zp) {
  return __fmul_rn(delta, __fsub_rn(code, zp));
}

// Same result as uaq_code, without the IEEE division on the common path: t = x * (1/delta) differs
// from the correctly rounded quotient by < 2^-22 |t|, so rint(t) is the reference's rint(x / delta)
// unless t lies within that distance of a rounding boundary (k + 1/2) -- only then is the exact
// division evaluated (~1e-4 of the elements).  rint through the 1.5 * 2^23 magic add (exact
// round-half-even for |t| < 2^22; larger |t| saturate the clamp anyway).
// N values at a time: the common path is branch-free (independent chains interleave), the rare exact
// path is ONE branch per batch.
__device__ __forceinline__ float uaq_round_rcp(float x, float inv_delta, bool& near_tie) {
  float t = __fmul_rn(x, inv_delta);
  t = fminf(fmaxf(t, -4.0e6f), 4.0e6f);
  const float r = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
  near_tie |= (0.5f - fabsf(__fsub_rn(t, r))) <= fabsf(t) * 4.76837158e-7f;
  return r;
}
template <int N>
__device__ __forceinline__ void uaq_codes_rcp(const float (&x)[N], const float (&delta)[N], const float (&inv)[N],
                                              const float (&zp)[N], float qmax, float (&code)[N]) {
  bool near_tie = false;
#pragma unroll
  for (int i = 0; i < N; ++i) code[i] = uaq_round_rcp(x[i], inv[i], near_tie);
  if (near_tie) {
#pragma unroll
    for (int i = 0; i < N; ++i) code[i] = rintf(__fdiv_rn(x[i], delta[i]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) code[i] = fminf(fmaxf(code[i] + zp[i], 0.0f), qmax);
}
// one (delta, zp) for the whole batch
template <int N>
__device__ __forceinline__ void uaq_codes_rcp1(const float (&x)[N], float delta, float inv, float zp, float qmax,
                                               float (&code)[N]) {
  bool near_tie = false;
#pragma unroll
  for (int i = 0; i < N; ++i) code[i] = uaq_round_rcp(x[i], inv, near_tie);
  if (near_tie) {
#pragma unroll
    for (int i = 0; i < N; ++i) code[i] = rintf(__fdiv_rn(x[i], delta));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) code[i] = fminf(fmaxf(code[i] + zp, 0.0f), qmax);
}

// rare paths kept out of line so the unrolled hot loops stay small (instruction cache)
static __device__ __noinline__ float uaq_round_exact(float x, float delta) { return rintf(__fdiv_rn(x, delta)); }
static __device__ __noinline__ float rcp_rn_slow(float d) { return __frcp_rn(d); }

// Lean form for the fused producers (no code output): value = delta * (code - zp), or the integer
// (code - zp) when kEmitInt.  ~12 FP32 instructions per element; `inv` must be the correctly rounded
// 1/delta.  No clamp before the magic-add rint: for |t| >= 2^22 the rounded value may be off by a few
// units, which either trips the near-tie test (exact path) or saturates the [0, qmax] clamp anyway.
template <bool kEmitInt, int N>
__device__ __forceinline__ void uaq_lean(float (&v)[N], const float (&delta)[N], const float (&inv)[N],
                                         const float (&zp)[N], float qmax) {
  float r[N];
  bool near_tie = false;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float t = __fmul_rn(v[i], inv[i]);
    r[i] = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
    near_tie |= fmaf(fabsf(t), -4.76837158e-7f, 0.5f - fabsf(__fsub_rn(t, r[i]))) <= 0.0f;
  }
  if (near_tie) {
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = uaq_round_exact(v[i], delta[i]);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float q = __fsub_rn(fminf(fmaxf(__fadd_rn(r[i], zp[i]), 0.0f), qmax), zp[i]);
    v[i] = kEmitInt ? q : __fmul_rn(delta[i], q);
  }
}
// one (delta, 1/delta, zp) for all N values
template <bool kEmitInt, int N>
__device__ __forceinline__ void uaq_lean1(float (&v)[N], float delta, float inv, float zp, float qmax) {
  float r[N];
  bool near_tie = false;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float t = __fmul_rn(v[i], inv);
    r[i] = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
    near_tie |= fmaf(fabsf(t), -4.76837158e-7f, 0.5f - fabsf(__fsub_rn(t, r[i]))) <= 0.0f;
  }
  if (near_tie) {
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = uaq_round_exact(v[i], delta);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float q = __fsub_rn(fminf(fmaxf(__fadd_rn(r[i], zp), 0.0f), qmax), zp);
    v[i] = kEmitInt ? q : __fmul_rn(delta, q);
  }
}

// uaq_lean with the clamp bounds folded: clamp(r + zp, 0, qmax) - zp == clamp(r, lo, hi) for lo = -zp,
// hi = qmax - zp (r and zp are integers: the sums are exact), and the near-tie test as one FMA + compare.
// 9 FP32 instructions per element instead of 12.
template <int N>
__device__ __forceinline__ void uaq_bounds(const float (&zp)[N], float qmax, float (&lo)[N], float (&hi)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) { lo[i] = -zp[i]; hi[i] = __fsub_rn(qmax, zp[i]); }
}
template <bool kEmitInt, int N>
__device__ __forceinline__ void uaq_lean_lh(float (&v)[N], const float (&delta)[N], const float (&inv)[N],
                                            const float (&lo)[N], const float (&hi)[N]) {
  float r[N];
  bool near_tie = false;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float t = __fmul_rn(v[i], inv[i]);
    r[i] = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
    near_tie |= fmaf(fabsf(t), 4.76837158e-7f, fabsf(__fsub_rn(t, r[i]))) >= 0.5f;
  }
  if (near_tie) {
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = uaq_round_exact(v[i], delta[i]);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float q = fminf(fmaxf(r[i], lo[i]), hi[i]);
    v[i] = kEmitInt ? q : __fmul_rn(delta[i], q);
  }
}
template <bool kEmitInt, int N>
__device__ __forceinline__ void uaq_lean1_lh(float (&v)[N], float delta, float inv, float lo, float hi) {
  float r[N];
  bool near_tie = false;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float t = __fmul_rn(v[i], inv);
    r[i] = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
    near_tie |= fmaf(fabsf(t), 4.76837158e-7f, fabsf(__fsub_rn(t, r[i]))) >= 0.5f;
  }
  if (near_tie) {
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = uaq_round_exact(v[i], delta);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float q = fminf(fmaxf(r[i], lo), hi);
    v[i] = kEmitInt ? q : __fmul_rn(delta, q);
  }
}

__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// erf-GELU (torch.nn.functional.gelu default), branch-free: erf by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7 absolute, the rounding level of fp32 erf itself), 2 MUFU + ~12 FP32 ops
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float y = x * 0.70710678118654752f;
  const float ay = fabsf(y);
  const float t = __frcp_rn(fmaf(0.3275911f, ay, 1.0f));
  float pl = fmaf(1.061405429f, t, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  pl *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ay * ay * -1.4426950408889634f));
  const float erfc_abs = pl * e;                              // erfc(|y|)
  const float cdf2 = y < 0.0f ? erfc_abs : 2.0f - erfc_abs;   // 1 + erf(y), no cancellation in the tail
  return 0.5f * x * cdf2;
}

union Half8 {
  uint4 u;
  __half2 h2[4];
  __half h[8];
};

__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  Half8 t;
  t.u = *reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(t.h2[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// fp16 pairs holding integers 0..255 -> their low bytes: h + 1024 = 0x6400 | value (exact), one PRMT gathers them
__device__ __forceinline__ uint32_t halves4_to_u8(uint32_t h01, uint32_t h23) {
  const __half2 k = __half2half2(__ushort_as_half(0x6400));
  const __half2 a = __hadd2(*reinterpret_cast<const __half2*>(&h01), k);
  const __half2 b = __hadd2(*reinterpret_cast<const __half2*>(&h23), k);
  return __byte_perm(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b), 0x6420);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  Half8 t;
#pragma unroll
  for (int i = 0; i < 4; ++i) t.h2[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return t.u;
}

}  // namespace dgq
