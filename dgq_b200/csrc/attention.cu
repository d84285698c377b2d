// Flash-style attention with DGQ's quantised softmax map, on tcgen05/TMEM.
//
// Replaces Attention.Attention_forward (diffusers_rewrite/sd.py:151-207, sdxl.py:174-229):
//   S = q_hat k_hat^T * d^-1/2 ; P = softmax(S) ; P_hat = aqtizer_w(P) ; O = P_hat v_hat
// with aqtizer_w = T2ILogQuantizer (quant/quant_layer_text.py:96-105; static or real-time delta,
// start-peak column bypass sd.py:191-195) or the always_zero UniformAffineQuantizer
// (quant/quant_block.py:149-156).  The (B,H,T,S) map never reaches HBM.
//
// The quantiser needs the FINAL softmax value p = exp(s - m_i) / l_i and, for real-time delta,
// the global max of the whole map, so the work is two passes over K:
//   pass 1  S = QK^T per tile -> row max m_i, row sum l_i (online), per-row max probability
//           -> atomicMax into gmax[0]                       (grid-wide dependency = kernel boundary)
//   pass 2  S recomputed -> codes -> P' (exact in fp16: 2^-code or the integer code) -> O += P' V
//           epilogue: O * delta (+ p_i0 * v_0 for the un-quantised start-peak column) -> fp16
// In base 2:  log2 p = s*alpha - beta_i,  alpha = scale*log2(e),  beta_i = M_i + log2 l_i.
//   log2 map : code = clamp(rint(beta_i + log2(delta) - s*alpha), 0, qmax),  P' = 2^-code
//   uniform  : code = clamp(rint(2^(s*alpha - beta_i - log2(delta))), 0, qmax), P' = code
//
// CTA = 128 query rows of one (batch, head); 192 threads:
//   warp 0 TMA loader | warp 1 MMA issuer (one lane) | warps 2..5 softmax: one thread per row
//   (TMEM lane == row, so row reductions need no shuffles).
// TMEM: S double-buffered (2 x 128 cols) + O (dp cols).  P' goes through smem in the UMMA
// K-major 128B-swizzle layout, written by the softmax threads.
#include "common.cuh"
#include "ptx.cuh"

namespace dgq {

constexpr int kAttThreads = 320;          // warp 0 loader, warp 1 MMA, warps 2..9 softmax
constexpr int kSoftmaxThreads = 256;
constexpr int kTileQ = 128;
constexpr int kTileK = 128;
constexpr uint32_t kChunkBytes = 128 * 64 * 2;  // one [128 x 64] fp16 SW128 sub-tile

struct AttnDev {
  int b, heads, t, s, d, dp;
  int nkv, kv_stages, q_tiles;
  float alpha;  // scale * log2(e)
  int map_mode, real_time, start_peak;
  const float* delta;
  float qmax;
  float* row_max;
  float* row_sum;
  float* gmax;
  const __half* vt;
  int sp;
  void* out;
  int ldo;
  int out_is_f32;
  uint8_t* codes;
  // quantizer of the consuming QuantLayer (to_out[0]) applied to O in the epilogue
  const float* oq_delta;
  const float* oq_zp;
  int oq_mode, oq_period, oq_emit_int;
  float oq_qmax;
};

// barrier indices
enum { B_QFULL = 0, B_KFULL = 1, B_KEMPTY = 3, B_VFULL = 5, B_VEMPTY = 7, B_SFULL = 9, B_SEMPTY = 11,
       B_PFULL = 13, B_PEMPTY = 15, B_OFULL = 17, B_COUNT = 18 };

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_bar_sync() {  // the 8 softmax warps only
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

// ---- pass 2, 32 scores of one row -> 32 fp16 operand values P' (packed in 16 regs)
//   MODE LOG2   : P' = 2^-code, code = clamp(rint(gamma - s*alpha), 0, qcap)   (no MUFU at all)
//   MODE UNIFORM: P' = code = min(rint(2^(s*alpha - gamma)), qmax)
//   MODE NONE   : P' = 2^(s*alpha - gamma)
template <int MODE, bool MASK, bool CODES>
__device__ __forceinline__ void map_chunk(const uint32_t (&r)[32], uint32_t (&h2)[16], float alpha, float gamma,
                                          float qcap, float qmax, int col0, int s_len, uint8_t* code_row) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float pv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float sc = __uint_as_float(r[i + e]);
      float val;
      if (MODE == DGQ_MAP_LOG2) {
        // rint through the 1.5*2^23 magic add; 2^-code rebuilt from the exponent field
        const float xq = fminf(fmaxf(fmaf(-alpha, sc, gamma), 0.f), qcap);
        const uint32_t yb = __float_as_uint(xq + 12582912.0f);
        val = __uint_as_float(yb * 0xFF800000u + 0x3F800000u);
      } else if (MODE == DGQ_MAP_UNIFORM) {
        val = fminf(rintf(ex2_approx(fmaf(alpha, sc, -gamma))), qmax);
      } else {
        val = ex2_approx(fmaf(alpha, sc, -gamma));
      }
      if (MASK) val = (col0 + i + e < s_len) ? val : 0.f;
      if (CODES) {
        if (col0 + i + e < s_len && MODE != DGQ_MAP_NONE) {
          const float cd = MODE == DGQ_MAP_LOG2 ? fminf(rintf(fmaxf(fmaf(-alpha, sc, gamma), 0.f)), qmax) : val;
          code_row[col0 + i + e] = static_cast<uint8_t>(cd);
        }
      }
      pv[e] = val;
    }
    const __half2 hh = __floats2half2_rn(pv[0], pv[1]);
    h2[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
  }
}

template <int PASS, int MODE, bool CODES>
__global__ void __launch_bounds__(kAttThreads, PASS == 1 ? 2 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int dchunks = p.dp >> 6;
  const uint32_t q_bytes = dchunks * kChunkBytes;          // Q tile / one K stage
  const uint32_t v_stage = 2 * p.dp * 128;                 // two [dp x 64] sub-tiles
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + q_bytes;
  uint8_t* s_v = s_k + p.kv_stages * q_bytes;
  uint8_t* s_p = s_v + (PASS == 2 ? p.kv_stages * v_stage : 0);
  uint8_t* s_end = s_p + (PASS == 2 ? 2 * 2 * kChunkBytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_end);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
  float* s_v0 = reinterpret_cast<float*>(tmem_slot + 2);   // [192] v_hat row 0 (start-peak)
  float* s_x = s_v0 + 192;                                 // [3][128] cross-half exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int nst = p.kv_stages;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_k);
    if (PASS == 2) prefetch_tmap(&tm_v);
    for (int i = 0; i < B_COUNT; ++i) {
      const bool all_sm = (i >= B_SEMPTY && i < B_SEMPTY + 2) || (i >= B_PFULL && i < B_PFULL + 2);
      mbar_init(&bars[i], all_sm ? 8 : 1);
    }
    fence_barrier_init();
  }
  constexpr uint32_t kTmemCols = PASS == 1 ? 256 : 512;   // pass 1 holds S only: two CTAs fit per SM
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (PASS == 2 && p.start_peak && threadIdx.x >= 64) {
    for (int dd = threadIdx.x - 64; dd < p.dp; dd += kSoftmaxThreads)
      s_v0[dd] = __half2float(p.vt[(static_cast<size_t>(bh) * p.dp + dd) * p.sp]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA loader
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[B_QFULL], q_bytes);
      for (int c = 0; c < dchunks; ++c)
        tma_load_3d(s_q + c * kChunkBytes, &tm_q, &bars[B_QFULL], c * 64, q_tile * kTileQ, bh);
      for (int j = 0; j < p.nkv; ++j) {
        const int slot = j % nst;
        const uint32_t ph = (j / nst) & 1;
        mbar_wait(&bars[B_KEMPTY + slot], ph ^ 1);
        mbar_arrive_expect_tx(&bars[B_KFULL + slot], q_bytes);
        for (int c = 0; c < dchunks; ++c)
          tma_load_3d(s_k + slot * q_bytes + c * kChunkBytes, &tm_k, &bars[B_KFULL + slot], c * 64, j * kTileK, bh);
        if (PASS == 2) {
          mbar_wait(&bars[B_VEMPTY + slot], ph ^ 1);
          mbar_arrive_expect_tx(&bars[B_VFULL + slot], v_stage);
          for (int c = 0; c < 2; ++c)
            tma_load_3d(s_v + slot * v_stage + c * (p.dp * 128), &tm_v, &bars[B_VFULL + slot], j * kTileK + c * 64, 0, bh);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(kTileQ, kTileK);
      const uint32_t idesc_o = umma_idesc_f16(kTileQ, p.dp);
      auto issue_pv = [&](int i) {
        const int slot = i % nst, pb = i & 1;
        mbar_wait(&bars[B_VFULL + slot], (i / nst) & 1);
        mbar_wait(&bars[B_PFULL + pb], (i >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint64_t da = umma_desc_sw128(smem_u32(s_p + pb * 2 * kChunkBytes + c * kChunkBytes));
          const uint64_t db = umma_desc_sw128(smem_u32(s_v + slot * v_stage + c * (p.dp * 128)));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_f16(tmem_o, da + 2 * ks, db + 2 * ks, idesc_o, (i | c | ks) != 0 ? 1u : 0u);
        }
        tc_commit(&bars[B_VEMPTY + slot]);
        tc_commit(&bars[B_PEMPTY + pb]);
      };
      mbar_wait(&bars[B_QFULL], 0);
      for (int j = 0; j < p.nkv; ++j) {
        const int slot = j % nst, sb = j & 1;
        mbar_wait(&bars[B_KFULL + slot], (j / nst) & 1);
        mbar_wait(&bars[B_SEMPTY + sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int c = 0; c < dchunks; ++c) {
          const uint64_t da = umma_desc_sw128(smem_u32(s_q + c * kChunkBytes));
          const uint64_t db = umma_desc_sw128(smem_u32(s_k + slot * q_bytes + c * kChunkBytes));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_f16(tmem_base + sb * kTileK, da + 2 * ks, db + 2 * ks, idesc_s, (c | ks) != 0 ? 1u : 0u);
        }
        tc_commit(&bars[B_KEMPTY + slot]);
        tc_commit(&bars[B_SFULL + sb]);
        if (PASS == 2 && j > 0) issue_pv(j - 1);
      }
      if (PASS == 2) {
        issue_pv(p.nkv - 1);
        tc_commit(&bars[B_OFULL]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps (2..9)
    // warp -> TMEM lane quarter (warp & 3) and column half ((warp - 2) >> 2): one thread per
    // (row, 64-column half) of the 128 x 128 score tile
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;           // row inside the tile == TMEM lane
    const int tq = q_tile * kTileQ + row;       // query index
    const bool row_ok = tq < p.t;
    const size_t ridx = static_cast<size_t>(bh) * p.t + tq;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const bool partial_last = (p.s % kTileK) != 0;

    if (PASS == 1) {
      float M = -INFINITY, Mx = -INFINITY, l = 0.f;
      for (int j = 0; j < p.nkv; ++j) {
        const int sb = j & 1;
        mbar_wait(&bars[B_SFULL + sb], (j >> 1) & 1);
        tc_fence_after();
        const bool mask = partial_last && j == p.nkv - 1;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = half * 2 + cc;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_addr + sb * kTileK + c * 32, r);
          tc_wait_ld();
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]) * p.alpha;
          if (mask) {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = (j * kTileK + c * 32 + i < p.s) ? x[i] : -INFINITY;
          }
          float cm = x[1];
#pragma unroll
          for (int i = 2; i < 32; ++i) cm = fmaxf(cm, x[i]);
          const float cmx = cm;                 // excludes element 0 of this chunk
          cm = fmaxf(cm, x[0]);
          Mx = fmaxf(Mx, (j == 0 && c == 0) ? cmx : cm);
          const float Mn = fmaxf(M, cm);
          if (Mn > -INFINITY) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += ex2_approx(x[i] - Mn);
            l = l * ex2_approx(M - Mn) + acc;
            M = Mn;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_SEMPTY + sb]);
      }
      // combine the two column halves of each row
      if (half == 1) { s_x[row] = M; s_x[128 + row] = l; s_x[256 + row] = Mx; }
      softmax_bar_sync();
      if (half == 0) {
        const float M1 = s_x[row], l1 = s_x[128 + row], Mx1 = s_x[256 + row];
        const float Mn = fmaxf(M, M1);
        if (Mn > -INFINITY) l = l * ex2_approx(M - Mn) + l1 * ex2_approx(M1 - Mn);
        M = Mn;
        Mx = fmaxf(Mx, Mx1);
        float rp = 0.f;
        if (row_ok) {
          p.row_max[ridx] = M;
          p.row_sum[ridx] = l;
          rp = (p.start_peak ? ex2_approx(Mx - M) : 1.0f) / l;
        }
        for (int o = 16; o > 0; o >>= 1) rp = fmaxf(rp, __shfl_xor_sync(0xffffffffu, rp, o));
        if (lane == 0 && p.real_time) atomicMax(reinterpret_cast<int*>(p.gmax), __float_as_int(rp));
      }
    } else {
      // ---------------------------------------------------------------- pass 2
      float delta = 1.0f;
      if (MODE != DGQ_MAP_NONE) delta = p.real_time ? p.gmax[0] : __ldg(p.delta);
      const float beta = row_ok ? (p.row_max[ridx] + log2f(p.row_sum[ridx])) : 0.f;
      const float gamma = beta + (MODE != DGQ_MAP_NONE ? log2f(delta) : 0.f);
      const float qcap = fminf(p.qmax, 126.f);
      float p0 = 0.f;                           // un-quantised start-peak probability of this row
      uint8_t* code_row = CODES ? p.codes + ridx * p.s : nullptr;
      for (int j = 0; j < p.nkv; ++j) {
        const int sb = j & 1;
        mbar_wait(&bars[B_SFULL + sb], (j >> 1) & 1);
        mbar_wait(&bars[B_PEMPTY + sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const bool mask = (partial_last && j == p.nkv - 1) || (CODES && !row_ok);
        uint8_t* sub = s_p + sb * 2 * kChunkBytes + half * kChunkBytes;   // this half's [128 x 64] sub-tile
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = half * 2 + cc;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_addr + sb * kTileK + c * 32, r);
          tc_wait_ld();
          uint32_t h2[16];
          const int col0 = j * kTileK + c * 32;
          if (mask) map_chunk<MODE, true, CODES>(r, h2, p.alpha, gamma, qcap, p.qmax, col0, row_ok ? p.s : 0, code_row);
          else map_chunk<MODE, false, CODES>(r, h2, p.alpha, gamma, qcap, p.qmax, col0, p.s, code_row);
          if (p.start_peak && j == 0 && c == 0) {
            p0 = ex2_approx(fmaf(p.alpha, __uint_as_float(r[0]), -beta));
            h2[0] &= 0xFFFF0000u;               // column 0 leaves the MMA; added back in the epilogue
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            *reinterpret_cast<uint4*>(sub + sw128_offset(row, cc * 4 + v)) =
                make_uint4(h2[4 * v], h2[4 * v + 1], h2[4 * v + 2], h2[4 * v + 3]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars[B_SEMPTY + sb]);
          mbar_arrive(&bars[B_PFULL + sb]);
        }
      }
      // ---- epilogue: O * out_scale (+ p0 * v0) -> out; the two halves split the dp columns
      if (p.start_peak) {
        if (half == 0) s_x[row] = p0;
        softmax_bar_sync();
        p0 = s_x[row];
      }
      mbar_wait(&bars[B_OFULL], 0);
      tc_fence_after();
      const float oscale = MODE == DGQ_MAP_NONE ? 1.0f : delta;
      const int head = bh % p.heads, bb = bh / p.heads;
      const size_t ooff = (static_cast<size_t>(bb) * p.t + tq) * p.ldo + head * p.d;
      __half* orow = static_cast<__half*>(p.out) + ooff;
      float* orow32 = static_cast<float*>(p.out) + ooff;
      const int dhalf = p.dp >> 1;
      for (int c = half * dhalf; c < (half + 1) * dhalf; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_o + lane_addr + c, r);
        tc_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int d0 = c + v * 8;
            if (d0 < p.d) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                f[i] = __uint_as_float(r[v * 8 + i]) * oscale;
                if (p.start_peak) f[i] = fmaf(p0, s_v0[d0 + i], f[i]);
              }
              if (p.oq_mode != DGQ_Q_NONE) {
                float qd[8], qz[8];
                if (p.oq_mode == DGQ_Q_KWISE) {
                  const int k0 = head * p.d + d0;
                  const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.oq_delta + k0));
                  const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.oq_delta + k0 + 4));
                  const float4 z0 = __ldg(reinterpret_cast<const float4*>(p.oq_zp + k0));
                  const float4 z1 = __ldg(reinterpret_cast<const float4*>(p.oq_zp + k0 + 4));
                  qd[0] = a0.x; qd[1] = a0.y; qd[2] = a0.z; qd[3] = a0.w; qd[4] = a1.x; qd[5] = a1.y; qd[6] = a1.z; qd[7] = a1.w;
                  qz[0] = z0.x; qz[1] = z0.y; qz[2] = z0.z; qz[3] = z0.w; qz[4] = z1.x; qz[5] = z1.y; qz[6] = z1.z; qz[7] = z1.w;
                } else {
                  const int j = p.oq_mode == DGQ_Q_ROWWISE ? static_cast<int>((static_cast<size_t>(bb) * p.t + tq) % p.oq_period) : 0;
                  const float dd = __ldg(p.oq_delta + j), zz = __ldg(p.oq_zp + j);
#pragma unroll
                  for (int i = 0; i < 8; ++i) { qd[i] = dd; qz[i] = zz; }
                }
                float qi[8], cd[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) qi[i] = __frcp_rn(qd[i]);
                uaq_codes_rcp<8>(f, qd, qi, qz, p.oq_qmax, cd);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  f[i] = p.oq_emit_int ? __fsub_rn(cd[i], qz[i]) : uaq_dequant(cd[i], qd[i], qz[i]);
              }
              if (p.out_is_f32) {
                *reinterpret_cast<float4*>(orow32 + d0) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(orow32 + d0 + 4) = make_float4(f[4], f[5], f[6], f[7]);
              } else {
                *reinterpret_cast<uint4*>(orow + d0) = pack8(f);
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();  // gemm.cu

// fp16 [batch, rows, cols] (contiguous) ; box = [1, box_rows, 64]
static int make_tmap_3d(CUtensorMap* map, const void* ptr, uint64_t batch, uint64_t rows, uint64_t cols,
                        uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return static_cast<int>(cudaErrorNotSupported);
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstride[2] = {cols * 2, rows * cols * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(cudaErrorInvalidValue);
}

}  // namespace dgq

extern "C" int dgq_attention(const dgq_attn_t* a, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(a != nullptr && a->q != nullptr && a->k != nullptr && a->vt != nullptr && a->out != nullptr);
  DGQ_CHECK_ARG(a->b > 0 && a->heads > 0 && a->t > 0 && a->s > 0 && a->d > 0);
  DGQ_CHECK_ARG(a->dp % 64 == 0 && a->dp >= a->d && a->dp <= 192 && a->d % 8 == 0);
  DGQ_CHECK_ARG(a->sp >= a->s && a->sp % 8 == 0 && a->ldo % 8 == 0);
  DGQ_CHECK_ARG(a->map_mode >= DGQ_MAP_NONE && a->map_mode <= DGQ_MAP_LOG2);
  DGQ_CHECK_ARG(a->row_max != nullptr && a->row_sum != nullptr && a->gmax != nullptr);
  DGQ_CHECK_ARG(a->map_mode == DGQ_MAP_NONE || a->real_time || a->delta != nullptr);
  DGQ_CHECK_ARG(!a->real_time || a->map_mode == DGQ_MAP_LOG2);
  DGQ_CHECK_ARG(a->out_q.mode >= DGQ_Q_NONE && a->out_q.mode <= DGQ_Q_ROWWISE);
  DGQ_CHECK_ARG(a->out_q.mode == DGQ_Q_NONE || (a->out_q.delta != nullptr && a->out_q.zp != nullptr));
  DGQ_CHECK_ARG(!(a->out_q.emit_int && a->out_q.mode == DGQ_Q_KWISE));

  AttnDev p;
  p.b = a->b; p.heads = a->heads; p.t = a->t; p.s = a->s; p.d = a->d; p.dp = a->dp;
  p.nkv = (a->s + kTileK - 1) / kTileK;
  p.kv_stages = a->dp <= 64 ? 2 : 1;   // smem: dp 128/192 leave room for one K/V stage only
  p.q_tiles = (a->t + kTileQ - 1) / kTileQ;
  p.alpha = a->scale * 1.4426950408889634f;
  p.map_mode = a->map_mode; p.real_time = a->real_time; p.start_peak = a->start_peak;
  p.delta = a->delta; p.qmax = a->qmax;
  p.row_max = a->row_max; p.row_sum = a->row_sum; p.gmax = a->gmax;
  p.vt = static_cast<const __half*>(a->vt); p.sp = a->sp;
  p.out = a->out; p.ldo = a->ldo; p.out_is_f32 = a->out_is_f32; p.codes = a->codes;
  p.oq_delta = a->out_q.delta; p.oq_zp = a->out_q.zp; p.oq_mode = a->out_q.mode;
  p.oq_period = a->out_q.period > 0 ? a->out_q.period : 1; p.oq_emit_int = a->out_q.emit_int;
  p.oq_qmax = a->out_q.qmax;

  const uint64_t bh = static_cast<uint64_t>(a->b) * a->heads;
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_3d(&tq, a->q, bh, a->t, a->dp, kTileQ);
  if (rc != 0) return rc;
  rc = make_tmap_3d(&tk, a->k, bh, a->s, a->dp, kTileK);
  if (rc != 0) return rc;
  rc = make_tmap_3d(&tv, a->vt, bh, a->dp, a->sp, a->dp);
  if (rc != 0) return rc;

  const uint32_t q_bytes = (a->dp / 64) * kChunkBytes;
  const uint32_t tail = 1024 + B_COUNT * 8 + 16 + 192 * 4 + 3 * 128 * 4 + 64;
  const uint32_t smem1 = q_bytes * (1 + p.kv_stages) + tail;
  const uint32_t smem2 = q_bytes * (1 + p.kv_stages) + p.kv_stages * 2 * a->dp * 128 + 4 * kChunkBytes + tail;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnDev);
  KernelFn k1 = attention_kernel<1, 0, false>;
  KernelFn k2;
  const bool cd = a->codes != nullptr;
  switch (a->map_mode) {
    case DGQ_MAP_LOG2: k2 = cd ? attention_kernel<2, DGQ_MAP_LOG2, true> : attention_kernel<2, DGQ_MAP_LOG2, false>; break;
    case DGQ_MAP_UNIFORM: k2 = cd ? attention_kernel<2, DGQ_MAP_UNIFORM, true> : attention_kernel<2, DGQ_MAP_UNIFORM, false>; break;
    default: k2 = attention_kernel<2, DGQ_MAP_NONE, false>; break;
  }
  // every instantiation gets the maximum it can ever need once (227 KB opt-in)
  static bool attr_done = false;
  if (!attr_done) {
    KernelFn all[] = {attention_kernel<1, 0, false>, attention_kernel<2, DGQ_MAP_LOG2, true>,
                      attention_kernel<2, DGQ_MAP_LOG2, false>, attention_kernel<2, DGQ_MAP_UNIFORM, true>,
                      attention_kernel<2, DGQ_MAP_UNIFORM, false>, attention_kernel<2, DGQ_MAP_NONE, false>};
    for (KernelFn f : all) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    attr_done = true;
  }
  if (smem2 > 232448) return DGQ_ERR_INVALID_VALUE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>(bh) * p.q_tiles;
  if (a->real_time) {
    cudaError_t e = cudaMemsetAsync(a->gmax, 0, sizeof(float), s);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  k1<<<grid, kAttThreads, smem1, s>>>(tq, tk, tv, p);
  k2<<<grid, kAttThreads, smem2, s>>>(tq, tk, tv, p);
  DGQ_RETURN_LAST_ERROR();
}
