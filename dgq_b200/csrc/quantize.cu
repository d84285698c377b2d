// Stand-alone quantizer kernels (HBM-bound, vectorised) and the offline weight packer.
//   dgq_fake_quant_f32    UniformAffineQuantizer.forward   quant/quant_layer.py:295-299
//   dgq_t2i_log_quant_f32 T2ILogQuantizer.forward          quant/quant_layer_text.py:96-105
//   dgq_max_f32           x.max() for real-time delta      quant/quant_layer_text.py:97
//   dgq_pack_weight       wqtizer(self.w) done once        quant/quant_layer.py:642-643,
//                         AdaRound hard rounding           quant/adaptive_rounding.py:51-70
#include "common.cuh"

namespace dgq {

// ------------------------------------------------------------------------------------------
__global__ void fake_quant_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ delta,
                                  const float* __restrict__ zp, int period, int64_t inner, float qmax,
                                  float* __restrict__ out_dq, uint8_t* __restrict__ out_codes) {
  const int64_t nvec = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + v);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    float dq[4];
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t e = v * 4 + i;
      const int j = period == 1 ? 0 : static_cast<int>((e / inner) % period);
      const float d = __ldg(delta + j), z = __ldg(zp + j);
      const float c = uaq_code(xs[i], d, z, qmax);
      dq[i] = uaq_dequant(c, d, z);
      packed |= static_cast<uint32_t>(c) << (8 * i);
    }
    if (out_dq != nullptr) reinterpret_cast<float4*>(out_dq)[v] = make_float4(dq[0], dq[1], dq[2], dq[3]);
    if (out_codes != nullptr) reinterpret_cast<uint32_t*>(out_codes)[v] = packed;
  }
  // tail (n % 4)
  const int64_t tail0 = nvec << 2;
  const int64_t t = tail0 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && t < n) {
    const int j = period == 1 ? 0 : static_cast<int>((t / inner) % period);
    const float d = delta[j], z = zp[j];
    const float c = uaq_code(x[t], d, z, qmax);
    if (out_dq != nullptr) out_dq[t] = uaq_dequant(c, d, z);
    if (out_codes != nullptr) out_codes[t] = static_cast<uint8_t>(c);
  }
}

// ------------------------------------------------------------------------------------------
// code = clamp(rint(-log2(x/delta)), 0, qmax); out = 2^-code * delta.  x = 0 -> +inf -> qmax.
__device__ __forceinline__ float t2i_code(float x, float delta, float qmax) {
  const float t = -log2f(__fdiv_rn(x, delta));
  return fminf(fmaxf(rintf(t), 0.0f), qmax);
}
__device__ __forceinline__ float t2i_dequant(float code, float delta) {
  // 2^-code built from its bit pattern: exact for normals and subnormals, 0 below 2^-149
  // (what torch's fp32 pow returns; codes >= 150 flush to zero)
  const int c = static_cast<int>(code);
  const float p = c <= 126 ? __int_as_float((127 - c) << 23) : (c <= 149 ? __int_as_float(1 << (149 - c)) : 0.0f);
  return __fmul_rn(p, delta);
}

__global__ void t2i_log_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ delta_p,
                               float qmax, float* __restrict__ out_dq, uint8_t* __restrict__ out_codes) {
  const float delta = __ldg(delta_p);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float c = t2i_code(__ldg(x + i), delta, qmax);
    if (out_dq != nullptr) out_dq[i] = t2i_dequant(c, delta);
    if (out_codes != nullptr) out_codes[i] = static_cast<uint8_t>(c);
  }
}

// ------------------------------------------------------------------------------------------
__global__ void max_partial_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partial) {
  float m = -INFINITY;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    m = fmaxf(m, __ldg(x + i));
  __shared__ float sm[32];
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
  }
}
__global__ void max_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, partial[i]);
  __shared__ float sm[32];
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) out[0] = m;
  }
}

// ------------------------------------------------------------------------------------------
// one thread per (n, tap, c) output element of the K-reordered operand
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ delta,
                                   const float* __restrict__ zp, const float* __restrict__ alpha, int n,
                                   int ci, int taps, int ci_pad, int n_pad, float qmax, int use_wq,
                                   uint8_t* __restrict__ codes, __half* __restrict__ operand) {
  const int k_out = taps * ci_pad;
  const int64_t total = static_cast<int64_t>(n_pad) * k_out;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int row = static_cast<int>(i / k_out);
    const int ko = static_cast<int>(i % k_out);
    const int tap = ko / ci_pad, c = ko % ci_pad;
    float op = 0.0f, code = 0.0f;
    if (row < n && c < ci) {
      const int64_t src = (static_cast<int64_t>(row) * ci + c) * taps + tap;  // [n][ci][taps]
      const float x = w[src];
      if (use_wq) {
        const float d = delta[row], z = zp[row];
        if (alpha != nullptr) {
          const float q = floorf(__fdiv_rn(x, d)) + (alpha[src] >= 0.0f ? 1.0f : 0.0f) + z;
          code = fminf(fmaxf(q, 0.0f), qmax);
        } else {
          code = uaq_code(x, d, z, qmax);
        }
        op = code - z;
      } else {
        op = x;
      }
    }
    if (codes != nullptr) codes[i] = static_cast<uint8_t>(code);
    operand[i] = __float2half_rn(op);
  }
}

__global__ void pack_nibbles_kernel(const uint8_t* __restrict__ codes, int64_t n_bytes, uint8_t* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_bytes; i += stride)
    out[i] = static_cast<uint8_t>((codes[2 * i] & 0xF) | (codes[2 * i + 1] << 4));
}

// compiled-checkpoint load path: integer codes (one per byte, or two 4-bit codes per byte, low nibble first)
// -> the resident fp16 operand (code - zp), zero in the padded rows / channels
__global__ void unpack_weight_kernel(const uint8_t* __restrict__ codes, int bits, const float* __restrict__ zp, int n,
                                     int ci, int taps, int ci_pad, int n_pad, __half* __restrict__ operand) {
  const int k_out = taps * ci_pad;
  const int64_t total = static_cast<int64_t>(n_pad) * k_out;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int row = static_cast<int>(i / k_out);
    const int c = static_cast<int>(i % k_out) % ci_pad;
    float op = 0.0f;
    if (row < n && c < ci) {
      const uint32_t code = bits == 4 ? ((codes[i >> 1] >> ((i & 1) * 4)) & 0xFu) : codes[i];
      op = static_cast<float>(code) - zp[row];
    }
    operand[i] = __float2half_rn(op);
  }
}

// weight codes -> s8 operand of the kind::i8 GEMM + per-row column-sum / offset tables (one warp per output channel)
__global__ void weight_to_i8_kernel(const uint8_t* __restrict__ codes, const float* __restrict__ zp, int n, int n_pad,
                                    int k_out, float qmax, int8_t* __restrict__ operand, int32_t* __restrict__ colsum,
                                    int32_t* __restrict__ b_off) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  const bool real = row < n;
  const int wz = real ? static_cast<int>(zp[row]) : 0;
  const int off = qmax <= 127.0f ? wz : 128;
  int sum = 0;
  for (int k = lane; k < k_out; k += 32) {
    const int v = real ? static_cast<int>(codes[static_cast<size_t>(row) * k_out + k]) - off : 0;
    operand[static_cast<size_t>(row) * k_out + k] = static_cast<int8_t>(v);
    sum += v;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) {
    colsum[row] = sum;
    if (b_off != nullptr) b_off[row] = real ? off - wz : 0;
  }
}

// one warp per output channel: per-tap sums of the s8 operand, folded into the 9 border classes
__global__ void conv_oob_colsum_kernel(const int8_t* __restrict__ operand, int n_pad, int c, int32_t* __restrict__ csoob) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  int tap_sum[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    int s = 0;
    for (int k = lane; k < c; k += 32) s += operand[static_cast<size_t>(row) * 9 * c + t * c + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    tap_sum[t] = s;
  }
  if (lane < 9) {
    const int cy = lane / 3, cx = lane % 3;       // class of the OUTPUT pixel: 0 = first row / column, 2 = last
    int s = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3 - 1, dx = t % 3 - 1;
      const bool out = (cy == 0 && dy < 0) || (cy == 2 && dy > 0) || (cx == 0 && dx < 0) || (cx == 2 && dx > 0);
      if (out) s += tap_sum[t];
    }
    csoob[lane * n_pad + row] = s;
  }
}

static int grid_for(int64_t work, int block, int max_blocks) {
  int64_t g = (work + block - 1) / block;
  if (g < 1) g = 1;
  return static_cast<int>(g < max_blocks ? g : max_blocks);
}

}  // namespace dgq

extern "C" int dgq_version(void) { return 100; }

extern "C" int dgq_fake_quant_f32(const float* x, int64_t n, const float* delta, const float* zp, int period,
                                  int64_t inner, float qmax, float* out_dq, uint8_t* out_codes, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(x != nullptr && delta != nullptr && zp != nullptr && n >= 0 && period >= 1 && inner >= 1);
  DGQ_CHECK_ARG(out_dq != nullptr || out_codes != nullptr);
  if (n == 0) return 0;
  const int grid = grid_for((n + 3) / 4, 256, kNumSMs * 8);
  fake_quant_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, delta, zp, period, inner, qmax,
                                                                        out_dq, out_codes);
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_t2i_log_quant_f32(const float* x, int64_t n, const float* delta, float qmax, float* out_dq,
                                     uint8_t* out_codes, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(x != nullptr && delta != nullptr && n >= 0);
  DGQ_CHECK_ARG(out_dq != nullptr || out_codes != nullptr);
  if (n == 0) return 0;
  t2i_log_kernel<<<grid_for(n, 256, kNumSMs * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, delta, qmax, out_dq, out_codes);
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_max_f32(const float* x, int64_t n, float* out, float* scratch, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(x != nullptr && out != nullptr && scratch != nullptr && n > 0);
  const int grid = grid_for(n, 256 * 4, 1024);
  max_partial_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, scratch);
  max_final_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(scratch, grid, out);
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_pack_weight(const float* w, const float* delta, const float* zp, const float* alpha, int n,
                               int ci, int taps, int ci_pad, int n_pad, float qmax, int use_wq, uint8_t* codes,
                               uint8_t* packed4, void* operand, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(w != nullptr && operand != nullptr && n > 0 && ci > 0 && taps > 0);
  DGQ_CHECK_ARG(ci_pad >= ci && n_pad >= n && (taps * ci_pad) % 8 == 0);
  DGQ_CHECK_ARG(!use_wq || (delta != nullptr && zp != nullptr));
  DGQ_CHECK_ARG(packed4 == nullptr || codes != nullptr);
  const int64_t total = static_cast<int64_t>(n_pad) * taps * ci_pad;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pack_weight_kernel<<<grid_for(total, 256, kNumSMs * 16), 256, 0, s>>>(
      w, delta, zp, alpha, n, ci, taps, ci_pad, n_pad, qmax, use_wq, codes, static_cast<__half*>(operand));
  if (packed4 != nullptr)
    pack_nibbles_kernel<<<grid_for(total / 2, 256, kNumSMs * 16), 256, 0, s>>>(codes, total / 2, packed4);
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_unpack_weight(const uint8_t* codes, int bits, const float* zp, int n, int ci, int taps, int ci_pad,
                                 int n_pad, void* operand, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(codes != nullptr && zp != nullptr && operand != nullptr);
  DGQ_CHECK_ARG((bits == 4 || bits == 8) && n > 0 && ci > 0 && taps > 0 && ci_pad >= ci && n_pad >= n);
  DGQ_CHECK_ARG(bits == 8 || (static_cast<int64_t>(taps) * ci_pad) % 2 == 0);
  const int64_t total = static_cast<int64_t>(n_pad) * taps * ci_pad;
  unpack_weight_kernel<<<grid_for(total, 256, kNumSMs * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      codes, bits, zp, n, ci, taps, ci_pad, n_pad, static_cast<__half*>(operand));
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_weight_to_i8(const uint8_t* codes, const float* zp, int n, int n_pad, int k_out, float qmax,
                                int8_t* operand, int32_t* colsum, int32_t* b_off, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(codes != nullptr && zp != nullptr && operand != nullptr && colsum != nullptr);
  DGQ_CHECK_ARG(n > 0 && n_pad >= n && k_out > 0 && qmax >= 1.0f && qmax <= 255.0f);
  weight_to_i8_kernel<<<(n_pad + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, zp, n, n_pad, k_out, qmax,
                                                                                   operand, colsum, b_off);
  DGQ_RETURN_LAST_ERROR();
}

extern "C" int dgq_conv_oob_colsum(const int8_t* operand, int n_pad, int c, int32_t* csoob, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(operand != nullptr && csoob != nullptr && n_pad > 0 && c > 0);
  conv_oob_colsum_kernel<<<(n_pad + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(operand, n_pad, c, csoob);
  DGQ_RETURN_LAST_ERROR();
}
