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
// Structure (one CTA per SM; 320 threads with the plain epilogue, 576 with the fused ones):
//   warp 0      TMA producer   : A tile [128 x 64] and B tile [bn x 64] per stage, 128-byte swizzle
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (M=128 | 256 as a CTA pair, N=bn, K=16) x 4 per stage
//   warps 2..   epilogue       : tcgen05.ld accumulator rows -> scale/bias/temb/resid (+GEGLU | head split, +quantize)
//                                -> shared-memory transpose -> coalesced stores (8 warps plain, 16 warps fused)
// Accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  Tiles are visited n-fastest so concurrently resident CTAs share A and B
// tiles through L2.
//
// kI8 (dgq_gemm_i8): the same pipeline on tcgen05.mma kind::i8 -- 2x the MMA rate (profiles/r2_probes.txt) -- for
// layers whose activation scale is constant along K (scalar / row-wise: every layer of the g=1 configs).
//   A = u8 activation codes, B = s8 (weight code - b_off[n]), s32 accumulate:
//   sum_k (c_mk - za_m)(w_nk - wz_n) = acc - za_m * colsum_n + e_n * (rowsum_m - K za_m),  e_n = b_off_n - wz_n
//   evaluated in INTEGER arithmetic in the epilogue (exact), then * delta_a[m] * delta_w[n] + bias as before.
//   colsum_n = sum_k B[n,k] is a pack-time table; rowsum_m = sum_k c_mk is needed only when e_n != 0 (W8: the
//   8-bit weight codes minus their zero point do not fit s8) and is computed in the kernel by 4 extra warps that
//   read each A stage from shared memory (dp4a), so no producer has to emit it.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace dgq {

constexpr int kBM = 128;           // rows of A per CTA
constexpr int kBK = 64;            // K elements per stage for fp16 operands (128 bytes); 128 for the u8 / s8 operands
constexpr int kMaxBN = 256;
// warp 0 TMA, warp 1 MMA, then the epilogue warps: 8 for the plain epilogue (2 column splits per TMEM
// lane quarter), 16 for the fused GEGLU / QKV epilogues (4 splits) -- those run ~40 dependent ALU
// instructions per result, so with 2 warps per scheduler they, not the MMAs, bounded the tile time
constexpr uint32_t kABytes = kBM * kBK * 2;        // 16 KB
constexpr int kStgLd = 36;                         // padded row stride (floats) of the fp32 epilogue transpose buffer
constexpr int kEpiTab = 9;                         // per-column tables staged per tile
constexpr int kRsRing = 16;                        // kI8: ring of per-tile row-sum vectors (the row-sum warp runs at most 2 + kStages tiles ahead)
enum { EPI_PLAIN = 0, EPI_GEGLU = 1, EPI_QKV = 2 };
template <int kEpi, bool kI8 = false, int kCtas = 2> struct EpiCfg {
  // (16 warps were measured for the kind::i8 plain epilogue too: slower -- it is bound by LSU wavefronts and
  // instruction issue, not by latency; profiles/r2_gemm_epilogue_digest.txt)
  static constexpr int kWarps = kEpi == EPI_PLAIN ? 8 : 16;
  static constexpr int kThreads = 64 + 32 * kWarps + (kI8 ? 32 : 0);   // kI8: + the row-sum warp
  static constexpr int kSplit = kWarps / 4;        // column splits of a tile (one per warp of a lane quarter)
  // per-warp transpose buffer: [32 rows][36] fp32 (plain), [32 rows][32] fp16 with a 16-byte XOR swizzle (fused)
  static constexpr uint32_t kStgBytes = kEpi == EPI_PLAIN ? 32 * kStgLd * 4 : 32 * 32 * 2;
};

struct EpiQuant {   // quantizer applied by the fused epilogues (the NEXT op's activation quantizer)
  const float* delta;
  const float* zp;
  int mode, period;
  float qmax;
  int emit_int;
};

// kSmall (single CTA only): N tile <= 64 columns and a 6-deep ring.  Small-M problems (batch 1: a handful of tiles,
// K up to 23040) are one long latency chain per CTA -- a k-block took ~0.55 us with 3 stages in flight whatever the
// tile width; narrower tiles put 4x the CTAs to work and the deeper ring doubles the bytes each keeps in flight.
constexpr int kSmallBN = 64;
template <int kCtas, int kEpi, bool kI8 = false, bool kSmall = false> struct GemmCfg {
  static constexpr int kStages = kSmall ? 6 : (kCtas == 1 ? 3 : 5);
  static constexpr uint32_t kBBytes = (kSmall ? kSmallBN : kMaxBN / kCtas) * kBK * 2;  // 32 KB, 16 KB per CTA of a pair, 8 KB small
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  // [2 buffers][scale | bias | q.delta | 1/q.delta | -q.zp | qmax - q.zp][256] fp32 + one [32 rows][36] fp32 transpose
  // buffer per epilogue warp
  static constexpr uint32_t kEpiBytes = 2 * kEpiTab * kMaxBN * 4 + EpiCfg<kEpi, kI8, kCtas>::kWarps * EpiCfg<kEpi, kI8, kCtas>::kStgBytes;
  static constexpr uint32_t kRsBytes = kI8 ? kRsRing * kBM * 4 : 0;
  static constexpr uint32_t kSmem = kStages * kStageBytes + kEpiBytes + kRsBytes + 1024 /*align*/ + 512 /*barriers*/;
};

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
  EpiQuant q2;
  // EPI_QKV geometry: GEMM row = (batch, token), column = (head, channel)
  int heads, d, dp, tokens, tp, transpose, skip_first;
  // EPI_QKV, K operand of the attention kernel: every scale of the score is folded into K (kfold[d] = the Q
  // quantizer's per-channel delta, or nullptr) and K is written as an fp16 hi | lo pair ([b, h, tokens, 2 dp]) so
  // that Q (integer codes - zp) . K carries ~22 bits: the softmax-map codes then follow the reference's to its own
  // fp32 rounding level (tests/test_layerwise_gpu.py)
  const float* kfold;
  int k_split;
  // kI8
  const int32_t* colsum;   // [n] sum_k B[n, k]
  const int32_t* b_off;    // [n] e_n = b_off_n - wz_n, or nullptr (= 0: the weight zero point is folded into B)
  const float* row_zp;     // activation zero point: row_zp[m % row_period]
  // kI8 + EPI_PLAIN, implicit 3x3 / stride 1 / pad 1 convolution: A is the NHWC u8 code tensor [cv_b, cv_h, cv_w, cv_c]
  // itself, read through a 4-D tensor map -- CTA tile = a (cv_bn x cv_bh x cv_bw) = 128-pixel patch, K block =
  // (tap, 128 channels), out-of-image taps zero-filled by TMA.  An exact-zero padding tap means code = zero point,
  // not code = 0, so the epilogue corrects per BORDER CLASS of the row (3 x 3: top / mid / bottom x left / mid /
  // right): colsum_n loses cv_csoob[class][n] = the column sums of the taps that fell outside, K loses their count.
  int cv_on, cv_b, cv_h, cv_w, cv_c, cv_bw, cv_bh, cv_bn, cv_cblocks, cv_lw, cv_lwh;
  const int32_t* cv_csoob;
  int cv_ldoob;
};

template <int kThreads> __device__ __forceinline__ void epi_bar_sync() {  // the epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
}

// kCtas == 1: one CTA per 128 x bn tile.  kCtas == 2: a CTA pair (cluster of 2, cta_group::2) per
// 256 x bn tile -- each CTA stages its own 128 rows of A and bn/2 rows of B, the leader issues the
// MMAs for both, each CTA drains its own 128 accumulator rows.  Per FLOP this moves 2/3 of the
// L2->smem bytes of the single-CTA tile.
template <int kCtas, int kEpi, bool kI8, bool kSmall = false>
__global__ void __launch_bounds__(EpiCfg<kEpi, kI8, kCtas>::kThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmDev p) {
  using Cfg = GemmCfg<kCtas, kEpi, kI8, kSmall>;
  constexpr int kBKe = kI8 ? 2 * kBK : kBK;      // K elements per 128-byte stage row
  constexpr int kStages = Cfg::kStages;
  constexpr int kEpiWarps = EpiCfg<kEpi, kI8, kCtas>::kWarps;
  constexpr int kSplit = EpiCfg<kEpi, kI8, kCtas>::kSplit;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ base: the pointer keeps its address space, so the
  // epilogue / softmax accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  float* s_epi = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kEpiBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator drained (leader's copy is the one used)
  uint64_t* mdone_bar = tempty_bar + 2;        // [kStages] kI8 pairs: the MMAs reading this stage have retired
  uint64_t* rfull_bar = mdone_bar + kStages;   // [kRsRing] kI8: row sums of a tile published
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull_bar + kRsRing);
  int* s_rowsum = reinterpret_cast<int*>(smem + kStages * Cfg::kStageBytes + Cfg::kEpiBytes + 512);   // [kRsRing][128]
  const bool need_rowsum = kI8 && p.b_off != nullptr;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kCtas == 2 ? cluster_ctarank() : 0u;
  const int worker = kCtas == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int workers = static_cast<int>(gridDim.x) / kCtas;
  const bool cv = kI8 && kEpi == EPI_PLAIN && p.cv_on != 0;
  const int k_blocks = cv ? 9 * p.cv_cblocks : (p.k + kBKe - 1) / kBKe;
  // implicit conv: pixel patch of CTA `rank` of tile row m_blk -> (first x, first y, first image)
  auto cv_origin = [&](int m_blk, int& px0, int& py0, int& b0) {
    const int pidx = m_blk * kCtas + static_cast<int>(rank);
    const int pw = p.cv_w >> p.cv_lw, ph = p.cv_h / p.cv_bh;
    px0 = (pidx % pw) << p.cv_lw;
    py0 = ((pidx / pw) % ph) * p.cv_bh;
    b0 = (pidx / (pw * ph)) * p.cv_bn;
  };
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int b_rows = p.bn / kCtas;             // rows of B staged by this CTA

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], need_rowsum ? 2 : 1);   // MMA commit (+ the row-sum warp)
      mbar_init(&mdone_bar[i], 1);
    }
    for (int i = 0; i < kRsRing; ++i) mbar_init(&rfull_bar[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps * kCtas);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kCtas == 2) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs of a pair)
    // All 32 lanes run the loop in lock-step (indices, barrier addresses and coordinates stay warp-uniform, on the
    // uniform datapath) and one elected lane issues -- as `if (lane == 0) { loop }` every descriptor lived in
    // per-thread registers and each tcgen05 / TMA instruction dragged a chain of R2UR moves (attention.cu, round 2).
    {
      uint32_t stage = 0, phase = 0;
      const uint32_t tx = (kABytes + static_cast<uint32_t>(b_rows) * kBK * 2) * kCtas;
      for (int tile = worker; tile < total_tiles; tile += workers) {
        const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
        const int row_a = m_blk * (kBM * kCtas) + static_cast<int>(rank) * kBM;
        const int row_b = n_blk * p.bn + static_cast<int>(rank) * b_rows;
        int px0 = 0, py0 = 0, b0 = 0;
        if (cv) cv_origin(m_blk, px0, py0, b0);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
          if (cv) {
            const int tap = kb / p.cv_cblocks, ka = (kb - tap * p.cv_cblocks) * kBKe;
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            if (kCtas == 1) {
              mbar_arrive_expect_tx(&full_bar[stage], tx);
              tma_load_4d(smem_a + stage * kABytes, &tmap_a, &full_bar[stage], ka, px0 + dx, py0 + dy, b0);
              tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], tap * p.cv_c + ka, row_b);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx);
              tma_load_4d_pair(smem_a + stage * kABytes, &tmap_a, &full_bar[stage], ka, px0 + dx, py0 + dy, b0);
              tma_load_2d_pair(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], tap * p.cv_c + ka, row_b);
            }
          } else if (kCtas == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], tx);
            tma_load_2d(smem_a + stage * kABytes, &tmap_a, &full_bar[stage], kb * kBKe, row_a);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kBKe, row_b);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx);
            tma_load_2d_pair(smem_a + stage * kABytes, &tmap_a, &full_bar[stage], kb * kBKe, row_a);
            tma_load_2d_pair(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kBKe, row_b);
          }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only; warp-converged, one
    // elected lane issues, operand descriptors precomputed on the uniform datapath)
    if (rank == 0) {
      const uint32_t idesc = kI8 ? umma_idesc_i8(kBM * kCtas, p.bn, false, true) : umma_idesc_f16(kBM * kCtas, p.bn);
      const uint64_t da0 = umma_desc_sw128(smem_u32(smem_a)), db0 = umma_desc_sw128(smem_u32(smem_b));
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = worker; tile < total_tiles; tile += workers) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kMaxBN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = da0 + stage * (kABytes >> 4), db = db0 + stage * (Cfg::kBBytes >> 4);
          if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks) {
            // advancing 16 halves / 32 bytes along K inside the swizzle atom: +2 in the >>4 address field
            if (kI8) {
              if (kCtas == 2) tc_mma_i8_pair(d_tmem, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) != 0 ? 1u : 0u);
              else tc_mma_i8(d_tmem, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) != 0 ? 1u : 0u);
            } else {
              if (kCtas == 2) tc_mma_f16_pair(d_tmem, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) != 0 ? 1u : 0u);
              else tc_mma_f16(d_tmem, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) != 0 ? 1u : 0u);
            }
          }
          // frees the smem stage (in both CTAs) once these MMAs retire
          if (kCtas == 2) tc_commit_pair(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);
          if (kI8 && kCtas == 2 && need_rowsum) tc_commit_pair(&mdone_bar[stage]);   // the row-sum warps' cue
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
          if (kCtas == 2) tc_commit_pair(&tfull_bar[acc]); else tc_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (kI8 && warp >= 2 + kEpiWarps) {
    // ------------------------------------------------------------ row sums of the u8 A tile (kI8, W8 weights only)
    // one warp; lane l owns rows l, l + 32, l + 64, l + 96 of this CTA's 128 x 128-byte A stage.  The 16-byte chunks
    // of a row are read in the rotated order (i + row) & 7 -- any order sums the same, and the 8 lanes of a
    // quarter-warp then touch 8 different bank groups
    if (need_rowsum) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = worker; tile < total_tiles; tile += workers, ++it) {
        uint32_t sum[4] = {0u, 0u, 0u, 0u};
        for (int kb = 0; kb < k_blocks; ++kb) {
          // single CTA: as soon as the stage has landed; pair: the transaction bytes are counted on the LEADER's
          // barrier only, so both CTAs take the multicast "MMAs of this stage retired" commit as their cue
          if (kCtas == 1) mbar_wait(&full_bar[stage], phase); else mbar_wait(&mdone_bar[stage], phase);
          const uint8_t* base = smem_a + stage * kABytes + lane * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 w = *reinterpret_cast<const uint4*>(base + q * 4096 + (((i + lane) & 7) << 4));
              sum[q] = __dp4a(w.x, 0x01010101u, sum[q]);
              sum[q] = __dp4a(w.y, 0x01010101u, sum[q]);
              sum[q] = __dp4a(w.z, 0x01010101u, sum[q]);
              sum[q] = __dp4a(w.w, 0x01010101u, sum[q]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // ring of kRsRing tiles: this warp can be at most 2 + kStages tiles ahead of the epilogue
#pragma unroll
        for (int q = 0; q < 4; ++q) s_rowsum[(it % kRsRing) * kBM + q * 32 + lane] = static_cast<int>(sum[q]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&rfull_bar[it % kRsRing]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // warp -> TMEM lane quarter (warp & 3) and column half.  Per 32-column chunk: tcgen05.ld (thread
    // = row) -> row_scale * scale[n] + bias[n] -> per-warp smem transpose -> lanes across columns:
    // + residual, 128-byte coalesced row-segment stores.  Per-column tables (scale, bias + the
    // time-embedding row when the tile lies inside one sample, the fused quantizer's delta / zp) are
    // staged in smem once per tile; residual segments are prefetched one chunk ahead.
    //   EPI_GEGLU: B rows are interleaved [32 x1 | 32 gate] per 64 columns (pack time); a chunk PAIR
    //              yields 32 features x1 * gelu(gate), quantised for ff.net.2, stored as its fp16 operand.
    //   EPI_QKV  : quantise with aqtizer_q/k/v and store head-split ([b,h,t,dp] or V^T [b,h,dp,tp]).
    const int quad = warp & 3;
    const int split = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;
    constexpr int kUnit = kEpi == EPI_GEGLU ? 64 : 32;   // accumulator columns consumed per iteration
    const int nch = p.bn / kUnit;
    const int c_begin = (nch * split + kSplit - 1) / kSplit;
    const int c_end = (nch * (split + 1) + kSplit - 1) / kSplit;
    const bool temb_tile = p.temb != nullptr && (p.rows_per_batch % (kBM * kCtas)) == 0;
    float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_epi + 2 * kEpiTab * kMaxBN) +
                                          (warp - 2) * EpiCfg<kEpi, kI8, kCtas>::kStgBytes);
    const int rl0 = lane >> 3;          // row (0..3) inside a group of 4 rows
    const int cq = (lane & 7) * 4;      // first of this lane's 4 columns inside the chunk
    const EpiQuant& q2 = p.q2;
    uint32_t acc = 0, acc_phase = 0, it = 0;
    for (int tile = worker; tile < total_tiles; tile += workers, ++it) {
      const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
      const int tile_row0 = m_blk * (kBM * kCtas) + static_cast<int>(rank) * kBM;
      const int warp_row0 = tile_row0 + quad * 32;
      int px0 = 0, py0 = 0, b0 = 0;
      if (cv) cv_origin(m_blk, px0, py0, b0);
      // implicit conv: row r of this CTA's tile -> output row (or -1 beyond the batch) and its border class
      auto cv_row = [&](int r, int& cls) -> int {
        const int x = px0 + (r & (p.cv_bw - 1)), y = py0 + ((r >> p.cv_lw) & (p.cv_bh - 1)), bb = b0 + (r >> p.cv_lwh);
        cls = (y == 0 ? 0 : (y == p.cv_h - 1 ? 6 : 3)) + (x == 0 ? 0 : (x == p.cv_w - 1 ? 2 : 1));
        return bb < p.cv_b ? (bb * p.cv_h + y) * p.cv_w + x : -1;
      };
      int my_cls = 4;
      const int row = cv ? cv_row(quad * 32 + lane, my_cls) : warp_row0 + lane;
      const bool row_ok = cv ? row >= 0 : row < p.m;
      const int ncol0 = n_blk * p.bn;
      float* s_scale = s_epi + (it & 1) * kEpiTab * kMaxBN;
      float* s_bias = s_scale + kMaxBN;
      float* s_qd = s_bias + kMaxBN;
      float* s_qi = s_qd + kMaxBN;
      float* s_qz = s_qi + kMaxBN;      // PLAIN-era name: holds lo = -zp (the fused epilogues clamp code - zp to [lo, hi])
      float* s_qh = s_qz + kMaxBN;      // hi = qmax - zp
      int* s_cs = reinterpret_cast<int*>(s_qh + kMaxBN);   // kI8: colsum_n
      int* s_eo = s_cs + kMaxBN;                            // kI8: e_n
      float* s_kf = reinterpret_cast<float*>(s_eo + kMaxBN);   // EPI_QKV: kfold of the column's channel
      for (int j = etid; j < p.bn; j += 32 * kEpiWarps) {
        const int n = ncol0 + j;
        float sc = 1.0f, bi = 0.0f;
        if (kI8) {
          s_cs[j] = n < p.n ? __ldg(p.colsum + n) : 0;
          s_eo[j] = (n < p.n && p.b_off != nullptr) ? __ldg(p.b_off + n) : 0;
        }
        if (n < p.n) {
          if (p.scale != nullptr) sc = __ldg(p.scale + n);
          if (p.bias != nullptr) bi = __ldg(p.bias + n);
          if (temb_tile && (cv ? b0 < p.cv_b : tile_row0 < p.m)) {
            const size_t off = static_cast<size_t>(cv ? b0 : tile_row0 / p.rows_per_batch) * p.ld_temb + n;
            bi += p.ep_is_f32 ? __ldg(static_cast<const float*>(p.temb) + off)
                              : __half2float(static_cast<const __half*>(p.temb)[off]);
          }
        }
        s_scale[j] = sc;
        s_bias[j] = bi;
        if (kEpi == EPI_QKV && p.kfold != nullptr) s_kf[j] = n < p.n ? __ldg(p.kfold + n % p.d) : 1.0f;
        if (kEpi != EPI_PLAIN && (q2.mode == DGQ_Q_KWISE || q2.mode == DGQ_Q_SCALAR)) {
          // table slot j: EPI_QKV column j of the tile (index = channel inside the head);
          //               EPI_GEGLU output feature j of the tile (bn / 2 of them)
          int qi = 0;
          bool ok = true;
          if (q2.mode == DGQ_Q_KWISE) {
            if (kEpi == EPI_QKV) { qi = n % p.d; ok = n < p.n; }
            else { qi = n_blk * (p.bn >> 1) + j; ok = j < (p.bn >> 1) && 2 * qi < p.n; }
          }
          const float dd = ok ? __ldg(q2.delta + qi) : 1.0f;
          s_qd[j] = dd;
          s_qi[j] = __frcp_rn(dd);
          const float zz = ok ? __ldg(q2.zp + qi) : 0.0f;
          s_qz[j] = -zz;
          s_qh[j] = __fsub_rn(q2.qmax, zz);
        }
      }
      const float rs = (p.row_scale != nullptr && row_ok) ? __ldg(p.row_scale + (row % p.row_period)) : 1.0f;
      // kI8: integer correction terms of this thread's row: -za and (rowsum - K * za)
      int nrz = 0, rsp = 0;
      if (kI8) {
        const int za = row_ok ? static_cast<int>(__ldg(p.row_zp + (row % p.row_period))) : 0;
        nrz = -za;
        if (need_rowsum) {
          mbar_wait(&rfull_bar[it % kRsRing], (it / kRsRing) & 1);
          rsp = s_rowsum[(it % kRsRing) * kBM + quad * 32 + lane] - p.k * za;
        }
      }
      // accumulator word j (tile column) -> float: kind::f16 holds fp32; kind::i8 holds s32 and the zero-point
      // corrections are applied in exact integer arithmetic first
      auto accf = [&](uint32_t raw, int j) -> float {
        if (kI8) return __int2float_rn(static_cast<int>(raw) + nrz * s_cs[j] + s_eo[j] * rsp);
        return __uint_as_float(raw);
      };
      // fused quantizer, row-indexed parameters (thread = row)
      float qd_row = 1.0f, qi_row = 1.0f, qz_row = 0.0f;
      bool q_on = kEpi != EPI_PLAIN && q2.mode != DGQ_Q_NONE;
      int tok = 0, bat = 0;
      int wbat = 0, wtok = 0;               // (batch, token) of the warp's first row
      if (kEpi == EPI_QKV) {
        wbat = warp_row0 / p.tokens; wtok = warp_row0 - wbat * p.tokens;
        bat = wbat; tok = wtok + lane;
        while (tok >= p.tokens) { tok -= p.tokens; ++bat; }
      }
      if (kEpi != EPI_PLAIN && q2.mode == DGQ_Q_ROWWISE && row_ok) {
        const int qi = kEpi == EPI_QKV ? max(tok - p.skip_first, 0) : row % q2.period;
        qd_row = __ldg(q2.delta + qi);
        qz_row = __ldg(q2.zp + qi);
        qi_row = __frcp_rn(qd_row);
      }
      const bool q_skip = kEpi == EPI_QKV && p.skip_first && tok == 0;   // start-peak: token 0 bypasses
      // quantise 32 results of this thread's row with the fused quantizer; table slots slot0 .. slot0 + 31
      // g[i] = acc[i] * row_scale * scale[n] + bias[n] for 32 consecutive tile columns j0 ..
      auto affine32 = [&](const uint32_t (&r)[32], int j0, float (&g)[32]) {
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + j0 + v * 4);
          const float4 bi = *reinterpret_cast<const float4*>(s_bias + j0 + v * 4);
          g[v * 4 + 0] = fmaf(accf(r[v * 4 + 0], j0 + v * 4 + 0) * rs, sc.x, bi.x);
          g[v * 4 + 1] = fmaf(accf(r[v * 4 + 1], j0 + v * 4 + 1) * rs, sc.y, bi.y);
          g[v * 4 + 2] = fmaf(accf(r[v * 4 + 2], j0 + v * 4 + 2) * rs, sc.z, bi.z);
          g[v * 4 + 3] = fmaf(accf(r[v * 4 + 3], j0 + v * 4 + 3) * rs, sc.w, bi.w);
        }
      };
      if constexpr (kEpi != EPI_PLAIN) {
        // ---- fused epilogues (16 warps: thread = one row x a quarter of the tile's columns).  A unit of 32
        // results is produced in two batches of 16 (register budget: 112 per thread at 576 threads), packed to
        // fp16, transposed through a swizzled 2 KB per-warp buffer and stored as 64-byte row segments.
        const bool has_rs = p.row_scale != nullptr;   // K-wise A operands carry their delta: no row scale
        auto affine16 = [&](const uint32_t (&r)[16], int j0, float (&g)[16]) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + j0 + v * 4);
            const float4 bi = *reinterpret_cast<const float4*>(s_bias + j0 + v * 4);
            if (has_rs || kI8) {
              g[v * 4 + 0] = fmaf(accf(r[v * 4 + 0], j0 + v * 4 + 0) * rs, sc.x, bi.x);
              g[v * 4 + 1] = fmaf(accf(r[v * 4 + 1], j0 + v * 4 + 1) * rs, sc.y, bi.y);
              g[v * 4 + 2] = fmaf(accf(r[v * 4 + 2], j0 + v * 4 + 2) * rs, sc.z, bi.z);
              g[v * 4 + 3] = fmaf(accf(r[v * 4 + 3], j0 + v * 4 + 3) * rs, sc.w, bi.w);
            } else {
              g[v * 4 + 0] = fmaf(__uint_as_float(r[v * 4 + 0]), sc.x, bi.x);
              g[v * 4 + 1] = fmaf(__uint_as_float(r[v * 4 + 1]), sc.y, bi.y);
              g[v * 4 + 2] = fmaf(__uint_as_float(r[v * 4 + 2]), sc.z, bi.z);
              g[v * 4 + 3] = fmaf(__uint_as_float(r[v * 4 + 3]), sc.w, bi.w);
            }
          }
        };
        const float lo_row = -qz_row, hi_row = __fsub_rn(q2.qmax, qz_row);
        const bool out_u8 = kEpi == EPI_GEGLU && q2.emit_int == 2;   // the operand of a kind::i8 consumer: u8 codes
        auto fused_quant16 = [&](float (&g)[16], int slot0) {
          if (!q_on || q_skip) return;
          if (q2.mode == DGQ_Q_ROWWISE) {
            if (q2.emit_int) uaq_lean1_lh<true, 16>(g, qd_row, qi_row, lo_row, hi_row);
            else uaq_lean1_lh<false, 16>(g, qd_row, qi_row, lo_row, hi_row);
            if (out_u8) {
#pragma unroll
              for (int i = 0; i < 16; ++i) g[i] += qz_row;          // (code - zp) + zp: the code, exact in fp16
            }
            return;
          }
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            float x[8], dd[8], ii[8], lo[8], hi[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = g[v * 8 + i];
#pragma unroll
            for (int h4 = 0; h4 < 2; ++h4) {
              const float4 a = *reinterpret_cast<const float4*>(s_qd + slot0 + v * 8 + h4 * 4);
              const float4 b = *reinterpret_cast<const float4*>(s_qi + slot0 + v * 8 + h4 * 4);
              const float4 c = *reinterpret_cast<const float4*>(s_qz + slot0 + v * 8 + h4 * 4);
              const float4 e = *reinterpret_cast<const float4*>(s_qh + slot0 + v * 8 + h4 * 4);
              dd[h4 * 4] = a.x; dd[h4 * 4 + 1] = a.y; dd[h4 * 4 + 2] = a.z; dd[h4 * 4 + 3] = a.w;
              ii[h4 * 4] = b.x; ii[h4 * 4 + 1] = b.y; ii[h4 * 4 + 2] = b.z; ii[h4 * 4 + 3] = b.w;
              lo[h4 * 4] = c.x; lo[h4 * 4 + 1] = c.y; lo[h4 * 4 + 2] = c.z; lo[h4 * 4 + 3] = c.w;
              hi[h4 * 4] = e.x; hi[h4 * 4 + 1] = e.y; hi[h4 * 4 + 2] = e.z; hi[h4 * 4 + 3] = e.w;
            }
            if (q2.emit_int) uaq_lean_lh<true, 8>(x, dd, ii, lo, hi);
            else uaq_lean_lh<false, 8>(x, dd, ii, lo, hi);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[v * 8 + i] = out_u8 ? x[i] - lo[i] : x[i];   // lo = -zp

          }
        };
        const uint32_t t_acc = tmem_base + acc * kMaxBN + (static_cast<uint32_t>(quad * 32) << 16);
        uint8_t* stg8 = reinterpret_cast<uint8_t*>(stg);
        const bool scatter_t = kEpi == EPI_QKV && p.transpose;
        const uint32_t ldk = (kEpi == EPI_QKV && p.k_split) ? 2u * p.dp : static_cast<uint32_t>(p.dp);   // K row stride
        epi_bar_sync<32 * kEpiWarps>();     // staged tables visible
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        for (int c = c_begin; c < c_end; ++c) {
          const int j0 = c * kUnit;         // first accumulator column of this unit inside the tile
          uint32_t hp[16];                  // 32 results of this thread's row, fp16 pairs
          uint32_t hl[kEpi == EPI_QKV ? 16 : 1];   // EPI_QKV k_split: their low parts
#pragma unroll
          for (int sb = 0; sb < 2; ++sb) {
            float g[16];
            uint32_t r1[16];
            tmem_ld_32x16(t_acc + j0 + sb * 16, r1);
            if (kEpi == EPI_GEGLU) {
              uint32_t r2[16];
              tmem_ld_32x16(t_acc + j0 + 32 + sb * 16, r2);
              tc_wait_ld();
              float x2[16];
              affine16(r1, j0 + sb * 16, g);
              affine16(r2, j0 + 32 + sb * 16, x2);
#pragma unroll
              for (int i = 0; i < 16; ++i) g[i] *= gelu_erf_f(x2[i]);
              fused_quant16(g, c * 32 + sb * 16);
            } else {
              tc_wait_ld();
              affine16(r1, j0 + sb * 16, g);
              fused_quant16(g, j0 + sb * 16);
              if (kEpi == EPI_QKV && p.kfold != nullptr) {   // K operand: the Q quantizer's per-channel delta
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                  const float4 f = *reinterpret_cast<const float4*>(s_kf + j0 + sb * 16 + v * 4);
                  g[v * 4 + 0] *= f.x; g[v * 4 + 1] *= f.y; g[v * 4 + 2] *= f.z; g[v * 4 + 3] *= f.w;
                }
              }
            }
            if (scatter_t) {
              // V^T [b, heads, dp, tp]: for a fixed column the 32 lanes hold 32 consecutive tokens
              if (row_ok) {
                int n = ncol0 + j0 + sb * 16;
                int hh = n / p.d, dd = n - hh * p.d;
                __half* base = p.out + (static_cast<size_t>(bat) * p.heads) * p.dp * p.tp + tok;
#pragma unroll
                for (int i = 0; i < 16; ++i, ++n) {
                  if (n < p.n) base[(static_cast<size_t>(hh) * p.dp + dd) * p.tp] = __float2half_rn(g[i]);
                  if (++dd == p.d) { dd = 0; ++hh; }
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const __half2 h = __floats2half2_rn(g[2 * i], g[2 * i + 1]);
                hp[sb * 8 + i] = *reinterpret_cast<const uint32_t*>(&h);
                if (kEpi == EPI_QKV && p.k_split) {          // lo = fp16(value - fp16(value)): the next 11 bits
                  const float2 hf = __half22float2(h);
                  const __half2 l = __floats2half2_rn(g[2 * i] - hf.x, g[2 * i + 1] - hf.y);
                  hl[sb * 8 + i] = *reinterpret_cast<const uint32_t*>(&l);
                }
              }
            }
          }
          const int n_parts = (kEpi == EPI_QKV && p.k_split && !scatter_t) ? 2 : 1;
          for (int part = 0; part < n_parts; ++part) {
          if (part == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) hp[i] = hl[i];
          }
          if (!scatter_t) {
            // row `lane`, 16-byte chunk v -> slot v ^ ((lane >> 1) & 3): conflict-free for the row-wise writes
            // and for the reads below (lane = (row & 3, 8-byte column group))
#pragma unroll
            for (int v = 0; v < 4; ++v)
              *reinterpret_cast<uint4*>(stg8 + lane * 64 + ((v ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(hp[4 * v], hp[4 * v + 1], hp[4 * v + 2], hp[4 * v + 3]);
            __syncwarp();
            // result column of this lane's 4 values: GEGLU feature index, otherwise the GEMM column
            const int n = kEpi == EPI_GEGLU ? n_blk * (p.bn >> 1) + c * 32 + cq : ncol0 + j0 + cq;
            const int n_lim = kEpi == EPI_GEGLU ? (p.n >> 1) : p.n;
            if (n < n_lim) {
              const int ch = (lane & 7) >> 1, sub = (lane & 1) * 8;
              // element offset of (row rl0 of the warp, column n); rows advance by 4: + 4 row strides, and in the
              // head-split layout + one batch stride - `tokens` row strides when a sample boundary is crossed
              // (offsets fit 32 bits: checked by the launcher)
              uint32_t o, row_step;
              int gt = 0;
              if (kEpi == EPI_QKV) {
                const int hh = n / p.d, dd = n - hh * p.d;
                int gb = wbat;
                gt = wtok + rl0;
                while (gt >= p.tokens) { gt -= p.tokens; ++gb; }
                o = (static_cast<uint32_t>(gb * p.heads + hh) * p.tokens + gt) * ldk + dd + part * p.dp;
                row_step = 4u * ldk;
              } else {
                o = static_cast<uint32_t>(warp_row0 + rl0) * p.ldc + n;
                row_step = 4u * p.ldc;
              }
              const uint32_t wrap_step = kEpi == EPI_QKV ? static_cast<uint32_t>(p.heads - 1) * p.tokens * ldk : 0u;
#pragma unroll
              for (int rr = 0; rr < 8; ++rr) {
                const int rl = rr * 4 + rl0;
                if (warp_row0 + rl < p.m) {
                  const uint2 x = *reinterpret_cast<const uint2*>(stg8 + rl * 64 + ((ch ^ ((rl >> 1) & 3)) << 4) + sub);
                  if (out_u8) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(p.out) + o) = halves4_to_u8(x.x, x.y);
                  else *reinterpret_cast<uint2*>(p.out + o) = x;
                }
                o += row_step;
                if (kEpi == EPI_QKV) {
                  gt += 4;
                  if (gt >= p.tokens) { gt -= p.tokens; o += wrap_step; }
                }
              }
            }
            __syncwarp();
          }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCtas == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      if constexpr (!kI8) {
      // ---- kind::f16 plain epilogue (round 1 form): the affine in phase A (thread = row), per-column tables from
      // shared memory, residual segments prefetched one chunk ahead; fits 168 registers without spilling.
      const size_t esz = p.ep_is_f32 ? 4 : 2;
      const char* temb_row = (p.temb != nullptr && !temb_tile && row_ok)
                                 ? static_cast<const char*>(p.temb) + static_cast<size_t>(row / p.rows_per_batch) * p.ld_temb * esz
                                 : nullptr;
      float4 t_cur[8], t_nxt[8];
      auto load_resid = [&](int c, float4 (&t)[8]) {
        const int n = ncol0 + c * 32 + cq;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int grow = warp_row0 + rr * 4 + rl0;
          float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow < p.m && n < p.n) {
            if (p.ep_is_f32) {
              u = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.resid) + static_cast<size_t>(grow) * p.ld_resid + n));
            } else {
              const uint2 raw = __ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(p.resid) + static_cast<size_t>(grow) * p.ld_resid + n));
              const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
              const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
              u = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
          }
          t[rr] = u;
        }
      };
      const bool has_resid = kEpi == EPI_PLAIN && p.resid != nullptr;
      const uint32_t t_row = tmem_base + acc * kMaxBN + (static_cast<uint32_t>(quad * 32) << 16);
      // ---- lean path of the fp32 residual stream (the hot case: every to_out / ff.net.2 / conv2 / proj_out
      // layer): full tile, fp32 result (+ fp32 residual), bias / temb already folded into the staged tables.
      // No per-element bounds or dtype branches; the next accumulator chunk is requested as soon as the
      // current one is in registers.
      if (kEpi == EPI_PLAIN && p.out == nullptr && (p.ep_is_f32 || (p.resid == nullptr && p.temb == nullptr)) &&
          temb_row == nullptr && tile_row0 + kBM <= p.m && ncol0 + p.bn <= p.n && c_begin < c_end) {
        const float* resid = static_cast<const float*>(p.resid);
        const size_t row_a = static_cast<size_t>(warp_row0 + rl0);
        auto ldres = [&](int c, float4 (&t)[8]) {
          const float* b = resid + row_a * p.ld_resid + ncol0 + c * 32 + cq;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) t[rr] = __ldg(reinterpret_cast<const float4*>(b + static_cast<size_t>(rr) * 4 * p.ld_resid));
        };
        if (has_resid) ldres(c_begin, t_cur);
        epi_bar_sync<32 * kEpiWarps>();                     // staged tables visible
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32(t_row + c_begin * 32, r);
        const bool has_rs = p.row_scale != nullptr;
        for (int c = c_begin; c < c_end; ++c) {
          const int j0 = c * 32;
          if (has_resid && c + 1 < c_end) ldres(c + 1, t_nxt);
          tc_wait_ld();
          float g[32];
          if (has_rs) {
            affine32(r, j0, g);
          } else {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const float4 sc = *reinterpret_cast<const float4*>(s_scale + j0 + v * 4);
              const float4 bi = *reinterpret_cast<const float4*>(s_bias + j0 + v * 4);
              g[v * 4 + 0] = fmaf(__uint_as_float(r[v * 4 + 0]), sc.x, bi.x);
              g[v * 4 + 1] = fmaf(__uint_as_float(r[v * 4 + 1]), sc.y, bi.y);
              g[v * 4 + 2] = fmaf(__uint_as_float(r[v * 4 + 2]), sc.z, bi.z);
              g[v * 4 + 3] = fmaf(__uint_as_float(r[v * 4 + 3]), sc.w, bi.w);
            }
          }
          if (c + 1 < c_end) tmem_ld_32x32(t_row + j0 + 32, r);
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<float4*>(stg + lane * kStgLd + v * 4) = make_float4(g[v * 4], g[v * 4 + 1], g[v * 4 + 2], g[v * 4 + 3]);
          __syncwarp();
          float* o = p.out_f32 + row_a * p.ldc + ncol0 + j0 + cq;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            float4 x = *reinterpret_cast<const float4*>(stg + (rr * 4 + rl0) * kStgLd + cq);
            if (has_resid) { x.x += t_cur[rr].x; x.y += t_cur[rr].y; x.z += t_cur[rr].z; x.w += t_cur[rr].w; }
            *reinterpret_cast<float4*>(o + static_cast<size_t>(rr) * 4 * p.ldc) = x;
          }
          __syncwarp();
          if (has_resid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) t_cur[i] = t_nxt[i];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCtas == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      // ---- general plain path: edge tiles, fp16 results, per-row time-embedding rows
      if (has_resid && c_begin < c_end) load_resid(c_begin, t_cur);
      epi_bar_sync<32 * kEpiWarps>();                       // staged tables visible
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t r[32];
        const int j0 = c * kUnit;           // first accumulator column of this iteration inside the tile
        tmem_ld_32x32(t_row + j0, r);
        if (has_resid && c + 1 < c_end) load_resid(c + 1, t_nxt);
        float g[32];                        // this thread's row, 32 result columns
        tc_wait_ld();
        affine32(r, j0, g);
        if (temb_row != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int n = ncol0 + j0 + i;
            if (n < p.n)
              g[i] += p.ep_is_f32 ? reinterpret_cast<const float*>(temb_row)[n]
                                  : __half2float(reinterpret_cast<const __half*>(temb_row)[n]);
          }
        }
#pragma unroll
        for (int v = 0; v < 8; ++v)
          *reinterpret_cast<float4*>(stg + lane * kStgLd + v * 4) = make_float4(g[v * 4], g[v * 4 + 1], g[v * 4 + 2], g[v * 4 + 3]);
        __syncwarp();
        const int n = ncol0 + j0 + cq;      // result column of this lane's 4 values
        if (n < p.n) {
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int rl = rr * 4 + rl0;
            const int grow = warp_row0 + rl;
            if (grow < p.m) {
              float4 x = *reinterpret_cast<const float4*>(stg + rl * kStgLd + cq);
              if (has_resid) { x.x += t_cur[rr].x; x.y += t_cur[rr].y; x.z += t_cur[rr].z; x.w += t_cur[rr].w; }
              const size_t o = static_cast<size_t>(grow) * p.ldc + n;
              if (p.out != nullptr) {
                const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
                *reinterpret_cast<uint2*>(p.out + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
              }
              if (p.out_f32 != nullptr) *reinterpret_cast<float4*>(p.out_f32 + o) = x;
            }
          }
        }
        __syncwarp();
        if (has_resid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) t_cur[i] = t_nxt[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCtas == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      continue;
      }
      // ---- plain epilogue.  Phase A (thread = row) only moves the raw accumulator chunk TMEM -> registers -> the
      // warp's padded smem tile.  Phase B (lane = 4 columns x every 4th row) does the arithmetic with ITS columns'
      // constants held in registers for the whole chunk and its 8 rows' constants for the whole tile -- the affine,
      // the kind::i8 integer zero-point corrections, + residual -- and stores 128-byte (fp32) / 64-byte (fp16) row
      // segments.  (Round 1 did the affine in phase A: every thread re-read every column's constants from shared
      // memory, 2-4 LDS per result; ncu showed the kind::i8 kernel bound there, profiles/r2_gemm_i8_vs_f16.txt.)
      const bool has_resid = p.resid != nullptr;
      const bool has_rs = p.row_scale != nullptr;
      const uint32_t t_row = tmem_base + acc * kMaxBN + (static_cast<uint32_t>(quad * 32) << 16);
      // row constants (activation delta, -zero point, row sum - K * zero point) live in the 4 pad columns of the
      // row's staging line: written once per tile by the row's own thread, read back by whichever lane finishes it
      {
        const int ri = ((has_rs || kI8) && row_ok) ? row % p.row_period : 0;
        const float ra = (has_rs && row_ok) ? __ldg(p.row_scale + ri) : 1.0f;
        int nz = 0, rp = 0;
        if (kI8) {
          const int za = row_ok ? static_cast<int>(__ldg(p.row_zp + ri)) : 0;
          nz = -za;
          // K of this row: implicit conv rows at the image border see fewer taps
          const int cy = my_cls / 3, cx = my_cls - cy * 3;
          const int kin = cv ? p.cv_c * ((cy == 1 ? 3 : 2) * (cx == 1 ? 3 : 2)) : p.k;
          if (need_rowsum) rp = s_rowsum[(it % kRsRing) * kBM + quad * 32 + lane] - kin * za;
        }
        *reinterpret_cast<uint4*>(stg + lane * kStgLd + 32) =
            make_uint4(__float_as_uint(ra), static_cast<uint32_t>(nz), static_cast<uint32_t>(rp), static_cast<uint32_t>(my_cls));
      }
      // (a variant that folds a scalar activation quantizer's -za * colsum_n and delta_a * delta_w[n] once per chunk
      // was measured on the epilogue-only probe: 314 us against 227 us -- dropped)
      auto plain = [&](auto full_tag) {
        constexpr bool kFullTile = decltype(full_tag)::value;   // interior tile: no row / column bounds
        float4 t_cur[8];
        auto ldres = [&](int c, float4 (&t)[8]) {
          const int n = ncol0 + c * 32 + cq;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            int cls_unused;
            const int grow = cv ? cv_row(quad * 32 + rr * 4 + rl0, cls_unused) : warp_row0 + rr * 4 + rl0;
            float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kFullTile || (grow >= 0 && grow < p.m && n < p.n)) {
              if (p.ep_is_f32) {
                u = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.resid) + static_cast<size_t>(grow) * p.ld_resid + n));
              } else {
                const uint2 raw = __ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(p.resid) + static_cast<size_t>(grow) * p.ld_resid + n));
                const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                u = make_float4(lo.x, lo.y, hi.x, hi.y);
              }
            }
            t[rr] = u;
          }
        };
        epi_bar_sync<32 * kEpiWarps>();                       // staged tables visible
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        uint32_t r[32];
        if (c_begin < c_end) tmem_ld_32x32(t_row + c_begin * 32, r);
        for (int c = c_begin; c < c_end; ++c) {
          const int j0 = c * 32;
          // requested before the accumulator wait: the latency is covered by phase A and the other epilogue warps
          if (has_resid) ldres(c, t_cur);
          tc_wait_ld();
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<uint4*>(stg + lane * kStgLd + v * 4) = make_uint4(r[v * 4], r[v * 4 + 1], r[v * 4 + 2], r[v * 4 + 3]);
          if (c + 1 < c_end) tmem_ld_32x32(t_row + j0 + 32, r);   // the next chunk travels while this one is finished
          __syncwarp();
          const int j = j0 + cq;                                  // this lane's 4 columns inside the tile
          const int n = ncol0 + j;
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + j);
          const float4 bi = *reinterpret_cast<const float4*>(s_bias + j);
          int4 cs = make_int4(0, 0, 0, 0), eo = make_int4(0, 0, 0, 0);
          if (kI8) {
            cs = *reinterpret_cast<const int4*>(s_cs + j);
            eo = *reinterpret_cast<const int4*>(s_eo + j);
          }
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int rl = rr * 4 + rl0;
            int cls = 4;
            const int grow = cv ? cv_row(quad * 32 + rl, cls) : warp_row0 + rl;
            const uint4 raw = *reinterpret_cast<const uint4*>(stg + rl * kStgLd + cq);
            float4 x;
            float ra = 1.0f;
            if (kI8) {      // exact integer zero-point corrections, then one conversion
              const uint4 rc = *reinterpret_cast<const uint4*>(stg + rl * kStgLd + 32);
              ra = __uint_as_float(rc.x);
              const int nz = static_cast<int>(rc.y), rp = static_cast<int>(rc.z);
              int4 c4 = cs;
              if (cv && cls != 4 && (kFullTile || n < p.n)) {   // border row: the out-of-image taps carry no weight
                const int4 ob = __ldg(reinterpret_cast<const int4*>(p.cv_csoob + static_cast<size_t>(cls) * p.cv_ldoob + n));
                c4.x -= ob.x; c4.y -= ob.y; c4.z -= ob.z; c4.w -= ob.w;
              }
              x.x = __int2float_rn(static_cast<int>(raw.x) + nz * c4.x + eo.x * rp);
              x.y = __int2float_rn(static_cast<int>(raw.y) + nz * c4.y + eo.y * rp);
              x.z = __int2float_rn(static_cast<int>(raw.z) + nz * c4.z + eo.z * rp);
              x.w = __int2float_rn(static_cast<int>(raw.w) + nz * c4.w + eo.w * rp);
            } else {
              if (has_rs) ra = stg[rl * kStgLd + 32];
              x = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w));
            }
            if (has_rs || kI8) {
              x.x = fmaf(x.x * ra, sc.x, bi.x); x.y = fmaf(x.y * ra, sc.y, bi.y);
              x.z = fmaf(x.z * ra, sc.z, bi.z); x.w = fmaf(x.w * ra, sc.w, bi.w);
            } else {
              x.x = fmaf(x.x, sc.x, bi.x); x.y = fmaf(x.y, sc.y, bi.y);
              x.z = fmaf(x.z, sc.z, bi.z); x.w = fmaf(x.w, sc.w, bi.w);
            }
            const bool ok = kFullTile || (grow >= 0 && grow < p.m && n < p.n);
            if (!kFullTile && p.temb != nullptr && !temb_tile && ok) {   // a tile that spans several samples
              const size_t off = static_cast<size_t>(grow / p.rows_per_batch) * p.ld_temb + n;
              if (p.ep_is_f32) {
                const float4 te = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.temb) + off));
                x.x += te.x; x.y += te.y; x.z += te.z; x.w += te.w;
              } else {
                const uint2 te = __ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(p.temb) + off));
                const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&te.x));
                const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&te.y));
                x.x += lo.x; x.y += lo.y; x.z += hi.x; x.w += hi.y;
              }
            }
            if (has_resid) { x.x += t_cur[rr].x; x.y += t_cur[rr].y; x.z += t_cur[rr].z; x.w += t_cur[rr].w; }
            if (ok) {
              const size_t o = static_cast<size_t>(grow) * p.ldc + n;
              if (p.out != nullptr) {
                const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
                *reinterpret_cast<uint2*>(p.out + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
              }
              if (p.out_f32 != nullptr) *reinterpret_cast<float4*>(p.out_f32 + o) = x;
            }
          }
          __syncwarp();
        }
      };
      // interior tiles with no per-row time-embedding rows take the branch-free instantiation
      const bool rows_full = cv ? b0 + p.cv_bn <= p.cv_b : tile_row0 + kBM <= p.m;
      if (rows_full && ncol0 + p.bn <= p.n && (p.temb == nullptr || temb_tile)) plain(std::true_type{});
      else plain(std::false_type{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCtas == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (kCtas == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
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

// row-major [rows, cols] with row stride ld (elements) of fp16 (esize 2) or u8 / s8 (esize 1);
// box = [box_rows, 128 bytes of columns], 128B swizzle
int make_tmap_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, int esize = 2) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return static_cast<int>(cudaErrorNotSupported);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * esize};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esize), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(cudaErrorInvalidValue);
}

// u8 NHWC [b, h, w, c] (contiguous); box = [bn, bh, bw, 128 channels]: 128 pixel rows of 128 bytes, 128B swizzle
int make_tmap_nhwc_u8(CUtensorMap* map, const void* ptr, uint64_t b, uint64_t h, uint64_t w, uint64_t c,
                      uint32_t bn, uint32_t bh, uint32_t bw) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return static_cast<int>(cudaErrorNotSupported);
  cuuint64_t gdim[4] = {c, w, h, b};
  cuuint64_t gstride[3] = {c, w * c, h * w * c};
  cuuint32_t box[4] = {128, bw, bh, bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(cudaErrorInvalidValue);
}

// N tile: the largest multiple of 32 (<= 256) that wastes the least padded work
static int pick_bn(int n, int step) {
  int best = step;
  double best_cost = 1e30;
  for (int bn = 256; bn >= step; bn -= step) {
    const int tiles = (n + bn - 1) / bn;
    // padded columns, with a mild penalty for narrow tiles (A re-read from smem per MMA)
    const double cost = static_cast<double>(tiles) * bn * (1.0 + 16.0 / bn);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

template <int kCtas, int kEpi, bool kI8, bool kSmall = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, cudaStream_t s) {
  static PerDeviceOnce attr;       // per instantiation, per device
  int dev;
  if (!attr.done(&dev)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel<kCtas, kEpi, kI8, kSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GemmCfg<kCtas, kEpi, kI8, kSmall>::kSmem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr.mark(dev);
  }
  const int tiles = p.m_tiles * p.n_tiles;
  constexpr int kGemmThreads = EpiCfg<kEpi, kI8, kCtas>::kThreads;
  if (kCtas == 1) {
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    gemm_f16_kernel<kCtas, kEpi, kI8, kSmall><<<grid, kGemmThreads, GemmCfg<kCtas, kEpi, kI8, kSmall>::kSmem, s>>>(ta, tb, p);
  } else {
    const int pairs = tiles < kNumSMs / 2 ? tiles : kNumSMs / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    cfg.blockDim = dim3(kGemmThreads, 1, 1);
    cfg.dynamicSmemBytes = GemmCfg<kCtas, kEpi, kI8, kSmall>::kSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_f16_kernel<kCtas, kEpi, kI8, kSmall>, ta, tb, p);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  DGQ_RETURN_LAST_ERROR();
}

}  // namespace dgq

static int gemm_dispatch(const dgq_gemm_t* a, void* stream, bool i8) {
  using namespace dgq;
  DGQ_CHECK_ARG(a != nullptr && a->a != nullptr && a->b != nullptr);
  DGQ_CHECK_ARG(a->m > 0 && a->n > 0 && a->k > 0);
  DGQ_CHECK_ARG(a->k % 8 == 0 && a->lda % 8 == 0 && a->ldb % 8 == 0 && a->n % 8 == 0 && a->ldc % 8 == 0);
  if (i8) {   // byte operands: 16-byte aligned rows; zero-point corrections need their tables
    DGQ_CHECK_ARG(a->k % 16 == 0 && a->lda % 16 == 0 && a->ldb % 16 == 0);
    DGQ_CHECK_ARG(a->colsum != nullptr && a->row_zp != nullptr && a->row_scale != nullptr);
  }
  // implicit 3x3 convolution: patch geometry (bw x bh x bn = 128 pixels)
  const bool conv = a->conv_h > 0;
  int cv_bw = 0, cv_bh = 0, cv_bn = 0;
  if (conv) {
    DGQ_CHECK_ARG(i8 && a->epi == DGQ_EPI_PLAIN && a->conv_csoob != nullptr && a->conv_ldoob >= a->n);
    DGQ_CHECK_ARG(a->conv_b > 0 && a->conv_w > 1 && a->conv_h > 1 && a->conv_c > 0 && a->conv_c % 16 == 0);
    DGQ_CHECK_ARG(a->k == 9 * a->conv_c && a->m == a->conv_b * a->conv_h * a->conv_w && a->row_period == 1);
    cv_bw = a->conv_w < 16 ? a->conv_w : 16;
    cv_bh = a->conv_h < 128 / cv_bw ? a->conv_h : 128 / cv_bw;
    cv_bn = 128 / (cv_bw * cv_bh);
    DGQ_CHECK_ARG((cv_bw & (cv_bw - 1)) == 0 && (cv_bh & (cv_bh - 1)) == 0 && cv_bw * cv_bh * cv_bn == 128);
    DGQ_CHECK_ARG(a->conv_w % cv_bw == 0 && a->conv_h % cv_bh == 0);
  }
  DGQ_CHECK_ARG(a->out != nullptr || a->out_f32 != nullptr);
  DGQ_CHECK_ARG(a->temb == nullptr || (a->rows_per_batch > 0 && a->ld_temb % 8 == 0));
  DGQ_CHECK_ARG(a->resid == nullptr || a->ld_resid % 8 == 0);
  DGQ_CHECK_ARG(a->epi >= DGQ_EPI_PLAIN && a->epi <= DGQ_EPI_QKV);
  if (a->epi != DGQ_EPI_PLAIN) {
    const dgq_quant_t& q = a->q2;
    DGQ_CHECK_ARG(a->out != nullptr && a->resid == nullptr && a->temb == nullptr);
    DGQ_CHECK_ARG(q.mode >= DGQ_Q_NONE && q.mode <= DGQ_Q_ROWWISE);
    DGQ_CHECK_ARG(q.mode == DGQ_Q_NONE || (q.delta != nullptr && q.zp != nullptr));
    DGQ_CHECK_ARG(q.mode != DGQ_Q_ROWWISE || q.period > 0);
    DGQ_CHECK_ARG(!(q.emit_int && q.mode == DGQ_Q_KWISE && a->epi != DGQ_EPI_QKV) && q.emit_int >= 0 && q.emit_int <= 2);
    DGQ_CHECK_ARG(q.emit_int != 2 || (a->epi == DGQ_EPI_GEGLU && q.mode != DGQ_Q_NONE));
  }
  if (a->epi == DGQ_EPI_GEGLU) DGQ_CHECK_ARG(a->n % 64 == 0 && a->ldc >= a->n / 2 && static_cast<int64_t>(a->m) * a->ldc < (int64_t(1) << 31));
  if (a->epi == DGQ_EPI_QKV) {
    DGQ_CHECK_ARG(a->heads > 0 && a->d > 0 && a->d % 8 == 0 && a->dp >= a->d && a->dp % 8 == 0);
    DGQ_CHECK_ARG(a->n == a->heads * a->d && a->tokens > 0 && a->m % a->tokens == 0);
    DGQ_CHECK_ARG(!a->transpose || (a->tp >= a->tokens && a->tp % 8 == 0));
    DGQ_CHECK_ARG(a->tokens >= 4 && static_cast<int64_t>(a->m) * a->heads * a->dp * (a->k_split ? 2 : 1) < (int64_t(1) << 31));
    DGQ_CHECK_ARG(!(a->k_split && a->transpose));
  }

  static int force_ctas = -1;   // DGQ_GEMM_CTAS=1|2 pins the variant (benchmarking); default: by problem size
  if (force_ctas < 0) {
    const char* env = getenv("DGQ_GEMM_CTAS");
    force_ctas = env != nullptr ? atoi(env) : 0;
  }
  GemmDev p;
  p.m = a->m; p.n = a->n; p.k = a->k;
  // fused epilogues split a tile's columns over 4 warps per lane quarter: widths that divide evenly
  p.bn = pick_bn(a->n, a->epi == DGQ_EPI_PLAIN ? 32 : (a->epi == DGQ_EPI_GEGLU ? 256 : 128));
  static int force_bn = -1;     // DGQ_GEMM_BN pins the N tile (benchmarking)
  if (force_bn < 0) {
    const char* env = getenv("DGQ_GEMM_BN");
    force_bn = env != nullptr ? atoi(env) : 0;
  }
  if (force_bn > 0 && force_bn <= kMaxBN && force_bn % (a->epi == DGQ_EPI_GEGLU ? 64 : 32) == 0) p.bn = force_bn;
  p.n_tiles = (a->n + p.bn - 1) / p.bn;
  // CTA pairs (256-row tiles) once there is enough work to fill the 74 pairs; small problems keep
  // 128-row tiles on single CTAs so more SMs get a tile
  int ctas = (a->m > kBM && ((a->m + 2 * kBM - 1) / (2 * kBM)) * p.n_tiles >= kNumSMs / 2) ? 2 : 1;
  if (force_ctas == 1 || force_ctas == 2) ctas = force_ctas;
  p.m_tiles = (a->m + kBM * ctas - 1) / (kBM * ctas);
  // few tiles and a long reduction (the batch-1 UNet: 1-10 tiles, K up to 23040): 64-column tiles on a 6-deep ring
  static int small_on = -1;     // DGQ_GEMM_SMALL=0 switches the variant off (A/B runs)
  if (small_on < 0) {
    const char* env = getenv("DGQ_GEMM_SMALL");
    small_on = (env == nullptr || atoi(env) != 0) ? 1 : 0;
  }
  const int kb_total = (a->k + (i8 ? 2 * kBK : kBK) - 1) / (i8 ? 2 * kBK : kBK);
  const bool small = small_on && ctas == 1 && force_bn <= 0 && p.bn > kSmallBN && p.m_tiles * p.n_tiles * 2 <= kNumSMs &&
                     kb_total >= 4 && a->n >= kSmallBN;
  if (small) {
    p.bn = kSmallBN;
    p.n_tiles = (a->n + p.bn - 1) / p.bn;
  }
  p.scale = a->scale; p.bias = a->bias;
  p.row_scale = a->row_scale; p.row_period = a->row_period > 0 ? a->row_period : 1;
  p.temb = a->temb;
  p.rows_per_batch = a->rows_per_batch; p.ld_temb = a->ld_temb;
  p.resid = a->resid; p.ld_resid = a->ld_resid;
  p.out = static_cast<__half*>(a->out); p.ldc = a->ldc; p.out_f32 = a->out_f32;
  p.ep_is_f32 = a->ep_is_f32;
  p.q2 = EpiQuant{a->q2.delta, a->q2.zp, a->q2.mode, a->q2.period > 0 ? a->q2.period : 1, a->q2.qmax, a->q2.emit_int};
  p.heads = a->heads; p.d = a->d > 0 ? a->d : 1; p.dp = a->dp; p.tokens = a->tokens > 0 ? a->tokens : 1;
  p.tp = a->tp; p.transpose = a->transpose; p.skip_first = a->skip_first;
  p.kfold = a->epi == DGQ_EPI_QKV ? a->kfold : nullptr; p.k_split = a->epi == DGQ_EPI_QKV ? a->k_split : 0;
  p.colsum = i8 ? a->colsum : nullptr; p.b_off = i8 ? a->b_off : nullptr; p.row_zp = i8 ? a->row_zp : nullptr;
  p.cv_on = conv ? 1 : 0;
  p.cv_b = a->conv_b; p.cv_h = a->conv_h; p.cv_w = a->conv_w; p.cv_c = a->conv_c;
  p.cv_bw = cv_bw; p.cv_bh = cv_bh; p.cv_bn = cv_bn;
  p.cv_cblocks = conv ? (a->conv_c + 127) / 128 : 0;
  p.cv_lw = 0; p.cv_lwh = 0;
  while (conv && (1 << p.cv_lw) < cv_bw) ++p.cv_lw;
  while (conv && (1 << p.cv_lwh) < cv_bw * cv_bh) ++p.cv_lwh;
  p.cv_csoob = a->conv_csoob; p.cv_ldoob = a->conv_ldoob;
  if (conv) {   // one CTA tile per 128-pixel patch (the last image group may be partial: rows beyond the batch are skipped)
    const int patches = (a->conv_w / cv_bw) * (a->conv_h / cv_bh) * ((a->conv_b + cv_bn - 1) / cv_bn);
    p.m_tiles = (patches + ctas - 1) / ctas;
  }

  CUtensorMap ta, tb;
  const int esize = i8 ? 1 : 2;
  int rc = conv ? make_tmap_nhwc_u8(&ta, a->a, a->conv_b, a->conv_h, a->conv_w, a->conv_c, cv_bn, cv_bh, cv_bw)
                : make_tmap_2d(&ta, a->a, a->m, a->k, a->lda, kBM, esize);
  if (rc != 0) return rc;
  // B rows beyond n are zero-filled by TMA (out-of-bounds box rows)
  rc = make_tmap_2d(&tb, a->b, a->n, a->k, a->ldb, p.bn / ctas, esize);
  if (rc != 0) return rc;

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (small) {
    if (i8) {
      if (a->epi == DGQ_EPI_GEGLU) return launch_gemm<1, EPI_GEGLU, true, true>(ta, tb, p, s);
      if (a->epi == DGQ_EPI_QKV) return launch_gemm<1, EPI_QKV, true, true>(ta, tb, p, s);
      return launch_gemm<1, EPI_PLAIN, true, true>(ta, tb, p, s);
    }
    if (a->epi == DGQ_EPI_GEGLU) return launch_gemm<1, EPI_GEGLU, false, true>(ta, tb, p, s);
    if (a->epi == DGQ_EPI_QKV) return launch_gemm<1, EPI_QKV, false, true>(ta, tb, p, s);
    return launch_gemm<1, EPI_PLAIN, false, true>(ta, tb, p, s);
  }
  if (i8) {
    if (ctas == 1) {
      if (a->epi == DGQ_EPI_GEGLU) return launch_gemm<1, EPI_GEGLU, true>(ta, tb, p, s);
      if (a->epi == DGQ_EPI_QKV) return launch_gemm<1, EPI_QKV, true>(ta, tb, p, s);
      return launch_gemm<1, EPI_PLAIN, true>(ta, tb, p, s);
    }
    if (a->epi == DGQ_EPI_GEGLU) return launch_gemm<2, EPI_GEGLU, true>(ta, tb, p, s);
    if (a->epi == DGQ_EPI_QKV) return launch_gemm<2, EPI_QKV, true>(ta, tb, p, s);
    return launch_gemm<2, EPI_PLAIN, true>(ta, tb, p, s);
  }
  if (ctas == 1) {
    if (a->epi == DGQ_EPI_GEGLU) return launch_gemm<1, EPI_GEGLU, false>(ta, tb, p, s);
    if (a->epi == DGQ_EPI_QKV) return launch_gemm<1, EPI_QKV, false>(ta, tb, p, s);
    return launch_gemm<1, EPI_PLAIN, false>(ta, tb, p, s);
  }
  if (a->epi == DGQ_EPI_GEGLU) return launch_gemm<2, EPI_GEGLU, false>(ta, tb, p, s);
  if (a->epi == DGQ_EPI_QKV) return launch_gemm<2, EPI_QKV, false>(ta, tb, p, s);
  return launch_gemm<2, EPI_PLAIN, false>(ta, tb, p, s);
}

extern "C" int dgq_gemm_f16(const dgq_gemm_t* a, void* stream) { return gemm_dispatch(a, stream, false); }
extern "C" int dgq_gemm_i8(const dgq_gemm_t* a, void* stream) { return gemm_dispatch(a, stream, true); }
