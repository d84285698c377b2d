// qGEMM for DGQ's QuantLayer on B200: persistent, warp-specialised tcgen05 kernel.
//
//   C[m, n] = (sum_k A[m, k] * B[n, k]) * scale[n] + bias[n] (+ temb[m / rows_per_batch, n]) (+ resid[m, n])
//
//   A = de-quantised activation x_hat = delta * (code - zp), fp16 (written by the producers)
//   B = integer weight (code - zp) held exactly in fp16 (dgq_pack_weight); scale[n] = weight delta
//
// Replaces F.linear / F.conv2d / W.view(Co,-1) @ x_unf of the reference
// (quant/quant_layer.py:649-659) together with the residual / time-embedding adds that follow
// (quant/quant_block.py:105-117, 165-186).
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      TMA producer   : A tile [128 x 64] and B tile [bn x 64] per stage, 128-byte swizzle
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (M=128, N=bn, K=16) x 4 per stage
//   warps 2..5  epilogue       : tcgen05.ld accumulator rows -> scale/bias/temb/resid -> fp16 stores
// Accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  Tiles are visited n-fastest so concurrently resident CTAs share A and B
// tiles through L2.
#include "common.cuh"
#include "ptx.cuh"

namespace dgq {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kMaxBN = 256;
constexpr int kStages = 4;
constexpr int kGemmThreads = 192;
constexpr uint32_t kABytes = kBM * kBK * 2;      // 16 KB
constexpr uint32_t kBBytes = kMaxBN * kBK * 2;   // 32 KB
constexpr uint32_t kGemmSmem = kStages * (kABytes + kBBytes) + 1024 /*align*/ + 256 /*barriers*/;

struct GemmDev {
  int m, n, k, bn;
  int m_tiles, n_tiles;
  const float* scale;
  const float* row_scale;
  int row_period;
  const float* bias;
  const void* temb;
  int rows_per_batch, ld_temb;
  const void* resid;
  int ld_resid;
  __half* out;
  int ldc;
  float* out_f32;
  int ep_is_f32;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * (kABytes + kBBytes));
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k_blocks = (p.k + kBK - 1) / kBK;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t tx = kABytes + static_cast<uint32_t>(p.bn) * kBK * 2;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          tma_load_2d(smem_a + stage * kABytes, &tmap_a, &full_bar[stage], kb * kBK, m_blk * kBM);
          tma_load_2d(smem_b + stage * kBBytes, &tmap_b, &full_bar[stage], kb * kBK, n_blk * p.bn);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(kBM, p.bn);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kMaxBN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * kABytes));
          const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * kBBytes));
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks) {
            // advancing 16 halves (32 B) along K inside the swizzle atom: +2 in the >>4 address field
            tc_mma_f16(d_tmem, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);      // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quarter this warp may read
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * kBM + quad * 32 + lane;
      const bool row_ok = row < p.m;
      const size_t esz = p.ep_is_f32 ? 4 : 2;
      const char* temb_row = (p.temb != nullptr && row_ok)
                                 ? static_cast<const char*>(p.temb) + static_cast<size_t>(row / p.rows_per_batch) * p.ld_temb * esz
                                 : nullptr;
      const char* resid_row = (p.resid != nullptr && row_ok)
                                  ? static_cast<const char*>(p.resid) + static_cast<size_t>(row) * p.ld_resid * esz
                                  : nullptr;
      const float rs = (p.row_scale != nullptr && row_ok) ? __ldg(p.row_scale + (row % p.row_period)) : 1.0f;
      const uint32_t t_row = tmem_base + acc * kMaxBN + (static_cast<uint32_t>(quad * 32) << 16);
      for (int c = 0; c < p.bn; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c, r);
        tc_wait_ld();
        const int n0 = n_blk * p.bn + c;
        if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int n = n0 + v * 8;
            if (n < p.n) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[v * 8 + i]) * rs;
              if (p.scale != nullptr) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + n));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.scale + n + 4));
                f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
                f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
              }
              if (p.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              }
              if (temb_row != nullptr) {
                float t[8];
                if (p.ep_is_f32) load8(reinterpret_cast<const float*>(temb_row) + n, t);
                else load8(reinterpret_cast<const __half*>(temb_row) + n, t);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += t[i];
              }
              if (resid_row != nullptr) {
                float t[8];
                if (p.ep_is_f32) load8(reinterpret_cast<const float*>(resid_row) + n, t);
                else load8(reinterpret_cast<const __half*>(resid_row) + n, t);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += t[i];
              }
              if (p.out != nullptr)
                *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(row) * p.ldc + n) = pack8(f);
              if (p.out_f32 != nullptr) {
                float* o = p.out_f32 + static_cast<size_t>(row) * p.ldc + n;
                *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle
int make_tmap_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return static_cast<int>(cudaErrorNotSupported);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(cudaErrorInvalidValue);
}

// N tile: the largest multiple of 32 (<= 256) that wastes the least padded work
static int pick_bn(int n) {
  int best = 32;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 32; bn -= 32) {
    const int tiles = (n + bn - 1) / bn;
    // padded columns, with a mild penalty for narrow tiles (A re-read from smem per MMA)
    const double cost = static_cast<double>(tiles) * bn * (1.0 + 16.0 / bn);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace dgq

extern "C" int dgq_gemm_f16(const dgq_gemm_t* a, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(a != nullptr && a->a != nullptr && a->b != nullptr);
  DGQ_CHECK_ARG(a->m > 0 && a->n > 0 && a->k > 0);
  DGQ_CHECK_ARG(a->k % 8 == 0 && a->lda % 8 == 0 && a->ldb % 8 == 0 && a->n % 8 == 0 && a->ldc % 8 == 0);
  DGQ_CHECK_ARG(a->out != nullptr || a->out_f32 != nullptr);
  DGQ_CHECK_ARG(a->temb == nullptr || (a->rows_per_batch > 0 && a->ld_temb % 8 == 0));
  DGQ_CHECK_ARG(a->resid == nullptr || a->ld_resid % 8 == 0);

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  GemmDev p;
  p.m = a->m; p.n = a->n; p.k = a->k;
  p.bn = pick_bn(a->n);
  p.m_tiles = (a->m + kBM - 1) / kBM;
  p.n_tiles = (a->n + p.bn - 1) / p.bn;
  p.scale = a->scale; p.bias = a->bias;
  p.row_scale = a->row_scale; p.row_period = a->row_period > 0 ? a->row_period : 1;
  p.temb = a->temb;
  p.rows_per_batch = a->rows_per_batch; p.ld_temb = a->ld_temb;
  p.resid = a->resid; p.ld_resid = a->ld_resid;
  p.out = static_cast<__half*>(a->out); p.ldc = a->ldc; p.out_f32 = a->out_f32;
  p.ep_is_f32 = a->ep_is_f32;

  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, a->a, a->m, a->k, a->lda, kBM);
  if (rc != 0) return rc;
  // B rows beyond n are zero-filled by TMA (out-of-bounds box rows)
  rc = make_tmap_2d(&tb, a->b, a->n, a->k, a->ldb, p.bn);
  if (rc != 0) return rc;

  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  gemm_f16_kernel<<<grid, kGemmThreads, kGemmSmem, static_cast<cudaStream_t>(stream)>>>(ta, tb, p);
  DGQ_RETURN_LAST_ERROR();
}
