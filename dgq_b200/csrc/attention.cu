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
// PERSISTENT CTAs: a work item is one or two 128-row query tiles ("halves") of one (batch, head); CTA c runs
// items c, c + grid, ... as ONE continuous pipeline -- the K/V rings, the S and P' buffers, the Q ring and the
// O accumulators all carry over from item to item, so the QK^T of the next item overlaps the softmax / PV /
// epilogue of the current one (cross-attention, S = 77, is a single K tile per item: without this it is launch-
// and latency-bound).
// Roles: warp 0 Q/K TMA loader | warp 1 QK^T issuer | pass 2: warp 2 V loader, warp 3 PV issuer |
//        softmax warps (8 in pass 1, 16 in pass 2; TMEM lane == row, so row reductions need no shuffles) |
//        long log2 self-attention (template TWO): one more warp, the QK^T issuer of the second query half.
//        Role warps run warp-converged (all lanes loop, one elected lane issues).
// TMEM: S 2-3 x 128 columns + O 1-2 x dp columns.  P' goes through smem in the UMMA K-major 128B-swizzle
// layout, written by the softmax threads -- or, with three S buffers (dp = 64), stays in tensor memory, written
// over its own score columns, as the PV MMA's A operand (ptm).
// Operands (round 2): Q = bare integers (code - zp), K = every scale folded in, as an fp16 hi | lo pair (two QK^T
// MMAs per tile): S is exact to fp32 rounding, see DESIGN.md 4.2.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dgq {
// role timeline for scripts/attn_trace.py (-DDGQ_ATTN_TRACE builds only): SM clock of CTA 0's role events per step
#ifdef DGQ_ATTN_TRACE
__device__ long long g_trace[8][512];
#define DGQ_TR(role, idx) do { if (PASS == 2 && blockIdx.x == 0 && static_cast<uint32_t>(idx) < 512u) g_trace[role][idx] = clock64(); } while (0)
#else
#define DGQ_TR(role, idx) do { } while (0)
#endif


// warp 0 loader, warp 1 MMA, then the softmax warps: 8 in pass 1 (thread = row x 64-column half; the pass
// is MUFU-bound and two CTAs share an SM), 16 in pass 2 (thread = row x 32-column quarter; one CTA per
// SM, and the ALU-only map needs 4 warps per scheduler to cover its issue latency)
template <int PASS, bool TWO = false> struct AttCfg {
  static constexpr int kSoftmaxWarps = PASS == 1 ? 8 : 16;
  static constexpr int kSoftmaxThreads = 32 * kSoftmaxWarps;
  // pass 2: warp 2 loads V on its own ring, so a K tile is never queued behind a V tile that waits for
  // a PV to retire (the K/V prefetch distance was what bounded the whole kernel); warp 3 issues the PV MMAs;
  // four role warps keep softmax warp w on TMEM lane quarter w & 3
  static constexpr int kFirstSoftmaxWarp = PASS == 1 ? 2 : 4;
  // TWO (ping-pong pass 2 of long self-attention): one more warp after the softmax warps -- the QK^T issuer of the
  // SECOND query half.  (672 threads leave 80 registers per thread instead of 96: short key sequences and the MUFU-heavy
  // uniform map lose more to that than the second issuer returns, so they keep one issuer.)  A role warp is one
  // thread of serial code (~10 cycles per instruction: barrier polls, descriptor arithmetic, 64 cycles per MMA it
  // feeds); issuing both halves' QK^T cost one warp ~4100 cycles per K tile against the ~1900 the softmax groups
  // need (clock64 timeline of CTA 0, scripts/attn_trace.py) -- it, not any pipe, bounded pass 2
  static constexpr int kQkWarpB = kFirstSoftmaxWarp + kSoftmaxWarps;
  static constexpr int kThreads = 32 * kFirstSoftmaxWarp + kSoftmaxThreads + (TWO ? 32 : 0);
  static constexpr int kSplit = kSoftmaxWarps / 4;     // column splits of the 128-wide score tile
  static constexpr int kCh = 4 / kSplit;               // 32-column chunks per thread
};
constexpr int kTileQ = 128;
constexpr int kTileK = 128;
constexpr uint32_t kChunkBytes = 128 * 64 * 2;  // one [128 x 64] fp16 SW128 sub-tile

struct AttnDev {
  int b, heads, t, s, d, dp;
  int nkv, q_tiles, items;
  int nh;                       // 128-row query halves per work item (2: each K/V tile is loaded once for both)
  int nq_buf, nk_buf, nv_buf, no_buf;
  int ns;                       // S buffers in tensor memory (2, or 3 for the ping-pong pass 2 at dp = 64)
  int ptm;                      // pass 2, ns == 3: P' stays in tensor memory (written over its own S tile) as the PV MMA's A operand
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
  // score scales: Q is the bare integer (code - zp) and K carries every K-side / per-channel scale (hi | lo split, see
  // dgq_gemm_t.k_split); q_scale[row % period] = the Q quantizer's scalar / per-token delta, applied through alpha
  const float* q_scale;
  int q_period;
  int k_split;
  const float* oq_delta;
  const float* oq_zp;
  const float* oq_inv;
  int oq_mode, oq_period, oq_emit_int;
  float oq_qmax;
};

// barrier indices (rings of up to 4)
constexpr int kRing = 4;
// S-full / S-empty come in two sets of three (ping-pong pass 2 with three S buffers and one QK^T issuer per query half):
// buffer sb is used by steps of BOTH halves in turn, and a waiter that sees only every second phase of a barrier
// cannot tell "two phases ago" from "now" by parity -- so each (buffer, half) pair has its own barrier: S-full[sb + 3 g]
// is committed by issuer g and awaited by softmax group g, S-empty[sb + 3 g] is awaited by issuer g and signalled by
// whoever releases the tile to it (the other half's step three earlier).  Everything else uses set 0.
enum { B_QFULL = 0, B_QEMPTY = B_QFULL + kRing, B_KFULL = B_QEMPTY + kRing, B_KEMPTY = B_KFULL + kRing,
       B_VFULL = B_KEMPTY + kRing, B_VEMPTY = B_VFULL + kRing, B_SFULL = B_VEMPTY + kRing, B_SEMPTY = B_SFULL + 6,
       B_PFULL = B_SEMPTY + 6, B_PEMPTY = B_PFULL + 3, B_OFULL = B_PEMPTY + 2, B_OEMPTY = B_OFULL + 2,
       B_COUNT = B_OEMPTY + 2 };

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= 0 on the FMA pipe (Cody-Waite split + degree-6 Taylor of 2^f on [-0.5, 0.5]: relative error
// 1.2e-7, the level of ex2.approx): pass 1 is bound by the 16-lane MUFU pipe (ncu: XU 80 % busy, issue 50 %),
// so one exponential in four is computed here instead (~11 FMA/ALU instructions) and the two pipes finish together.
constexpr int kPolyEvery = 6;   // pass 1: one exponential in kPolyEvery goes to the FMA pipe
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);                       // masked scores are -inf
  const float t = x + 12582912.0f;             // integer part in the low mantissa bits
  const float f = x - (t - 12582912.0f);       // [-0.5, 0.5]
  float p = 1.5403530e-4f;
  p = fmaf(p, f, 1.33335581e-3f);
  p = fmaf(p, f, 9.61812911e-3f);
  p = fmaf(p, f, 5.550410866e-2f);
  p = fmaf(p, f, 2.4022650696e-1f);
  p = fmaf(p, f, 6.9314718056e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
template <int kThreads> __device__ __forceinline__ void softmax_bar_sync(int id = 1) {  // the softmax warps only
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kThreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- pass 2, 32 scores of one row -> 32 fp16 operand values P' (packed in 16 regs)
//   MODE LOG2   : P' = 2^-code, code = clamp(rint(gamma - s*alpha), 0, qcap)   (no MUFU at all)
//   MODE UNIFORM: P' = code = min(rint(2^(s*alpha - gamma)), qmax)
//   MODE NONE   : P' = 2^(s*alpha - gamma)
// Columns beyond the key length need no mask for LOG2 / UNIFORM: their P' is finite and the matching
// V^T columns are zero (TMA out-of-bounds fill / zeroed padding).  MODE NONE masks (2^-gamma may overflow).
template <int MODE, bool CODES>
__device__ __forceinline__ void map_chunk(const uint32_t (&r)[32], uint32_t (&h2)[16], float alpha, float gamma,
                                          float qcap, float qmax, int col0, int s_len, uint8_t* code_row) {
  // qcap is a POWER OF TWO (64, see the caller): s * (-alpha / qcap) + gamma / qcap is then the same rounding as
  // gamma - s * alpha with the exponent shifted, and y * qcap is exact -- the code below IS rint(gamma - s * alpha).
  // (Round 1 used qcap = 126: two roundings, ~7e-6 of a code step, i.e. code flips at ~1.5e-5 that the verification
  // output -- computed by the direct formula -- did not even show.)
  const float q_inv = __frcp_rn(qcap);
  const float a_sat = -alpha * q_inv, g_sat = gamma * q_inv;   // loop-invariant: hoisted by the compiler
  if (MODE == DGQ_MAP_LOG2 && !CODES && qmax >= qcap) {
    // the hot case as straight-line code (the level test inside the unrolled loop compiled to a branch per pair)
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float y0, y1;
      asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(y0) : "f"(__uint_as_float(r[i])), "f"(a_sat), "f"(g_sat));
      asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(y1) : "f"(__uint_as_float(r[i + 1])), "f"(a_sat), "f"(g_sat));
      const uint32_t b0 = __float_as_uint(fmaf(y0, qcap, 12582912.0f)), b1 = __float_as_uint(fmaf(y1, qcap, 12582912.0f));
      const __half2 hh = __floats2half2_rn(__uint_as_float(b0 * 0xFF800000u + 0x3F800000u),
                                           __uint_as_float(b1 * 0xFF800000u + 0x3F800000u));
      h2[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float pv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float sc = __uint_as_float(r[i + e]);
      float val, cd = 0.f;
      if (MODE == DGQ_MAP_LOG2) {
        if (qmax >= qcap) {
          // x / qcap clamped to [0, 1] by the FMA's own .sat (the two FMNMX of an explicit clamp run on the
          // half-rate ALU pipe, which bounded this loop); rint(x) through the 1.5 * 2^23 magic add of a second FMA;
          // 2^-code rebuilt from the exponent field by one IMAD.  Codes beyond qcap (>= 25 already flush to 0 in the
          // fp16 operand) stay at qcap: 2^-64.
          float y;
          asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(sc), "f"(a_sat), "f"(g_sat));
          const uint32_t yb = __float_as_uint(fmaf(y, qcap, 12582912.0f));
          val = __uint_as_float(yb * 0xFF800000u + 0x3F800000u);
          if (CODES) {
            cd = static_cast<float>(yb - 0x4B400000u);      // the code the operand was built from ...
            if (cd >= qcap) cd = fminf(rintf(fmaxf(fmaf(-alpha, sc, gamma), 0.f)), qmax);   // ... or beyond the cap
          }
        } else {
          // fewer levels than the cap (softmax_a_bit <= 4): explicit clamp at qmax
          cd = fminf(rintf(fmaxf(fmaf(-alpha, sc, gamma), 0.f)), qmax);
          val = __uint_as_float(0x3F800000u - (static_cast<uint32_t>(cd) << 23));
        }
      } else if (MODE == DGQ_MAP_UNIFORM) {
        val = fminf(rintf(ex2_approx(fmaf(alpha, sc, -gamma))), qmax);
        cd = val;
      } else {
        val = (col0 + i + e < s_len) ? ex2_approx(fmaf(alpha, sc, -gamma)) : 0.f;
      }
      if (CODES) {
        if (col0 + i + e < s_len && MODE != DGQ_MAP_NONE) code_row[col0 + i + e] = static_cast<uint8_t>(cd);
      }
      pv[e] = val;
    }
    const __half2 hh = __floats2half2_rn(pv[0], pv[1]);
    h2[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
  }
}

// PP ("ping-pong", pass 2 with two query halves per item): softmax warps 0..7 own query half 0 (S / P' buffer 0,
// O accumulator 0), warps 8..15 own half 1; thread = (row, 64-column half of the score tile).  The two groups
// run half a step apart, so one group's TMEM-load / proxy-fence / barrier latencies and its whole O epilogue
// are covered by the other group's ALU work (with all 16 warps in lock-step on one tile they were not:
// 45 % issue utilisation, ncu profiles/r1e_attention.txt).
template <int PASS, int MODE, bool CODES, bool PP, bool TWO = false>
__global__ void __launch_bounds__(AttCfg<PASS, TWO>::kThreads, PASS == 1 ? 2 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const AttnDev p) {
  using Cfg = AttCfg<PASS, TWO>;
  constexpr int kSoftmaxWarps = Cfg::kSoftmaxWarps, kSoftmaxThreads = Cfg::kSoftmaxThreads;
  constexpr int kSplit = PP ? 2 : Cfg::kSplit, kCh = PP ? 2 : Cfg::kCh;
  constexpr int kGroupThreads = PP ? kSoftmaxThreads / 2 : kSoftmaxThreads;   // threads sharing one score tile
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ base: the pointer keeps its address space, so the
  // epilogue / softmax accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int dchunks = p.dp >> 6;
  const uint32_t q_bytes = dchunks * kChunkBytes;          // one Q tile
  const int kparts = p.k_split ? 2 : 1;                    // a K tile travels as `kparts` ring entries: hi, then lo
  const uint32_t v_stage = 2 * p.dp * 128;                 // two [dp x 64] sub-tiles
  const uint32_t nqb = p.nq_buf, nkb = p.nk_buf, nvb = p.nv_buf, nob = p.no_buf, nh = p.nh;
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + nqb * q_bytes;
  uint8_t* s_v = s_k + nkb * q_bytes;
  uint8_t* s_p = s_v + (PASS == 2 ? nvb * v_stage : 0);
  uint8_t* s_end = s_p + (PASS == 2 ? 2 * 2 * kChunkBytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_end);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
  float* s_v0 = reinterpret_cast<float*>(tmem_slot + 2);   // [2 groups][2][192] v_hat row 0 of the item's head (start-peak)
  float* s_x = s_v0 + 4 * 192;                             // [3][128] cross-split exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_k);
    if (PASS == 2) prefetch_tmap(&tm_v);
    for (int i = 0; i < B_COUNT; ++i) {
      // (P' in tensor memory: the S tile is released by the PV MMAs' commit, not by the softmax warps)
      const bool all_sm = (i >= B_SEMPTY && i < B_SEMPTY + 6 && !(PASS == 2 && p.ptm)) || (i >= B_PFULL && i < B_PFULL + 3) ||
                          (i >= B_OEMPTY && i < B_OEMPTY + 2);
      const bool k_two = TWO && i >= B_KEMPTY && i < B_KEMPTY + kRing;   // one commit per QK^T issuer
      mbar_init(&bars[i], all_sm ? (PP ? kSoftmaxWarps / 2 : kSoftmaxWarps) : (k_two ? 2 : 1));
    }
    fence_barrier_init();
  }
  constexpr uint32_t kTmemCols = PASS == 1 ? 256 : 512;   // pass 1 holds S only: two CTAs fit per SM
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // score buffers: step U uses S[U % ns].  ns = 3 (ping-pong pass 2, dp = 64: 3 x 128 + 2 x 64 = 512 columns) lets the
  // QK^T of a group's NEXT tile start before that group has drained the current one -- with one buffer per group the
  // issuer waited for SEMPTY half of the time and the softmax warps for SFULL a quarter of theirs (ncu, round 2)
  const uint32_t ns = p.ns;
  const uint32_t tmem_o = tmem_base + ns * kTileK;
  const int q_groups = (p.q_tiles + static_cast<int>(nh) - 1) / static_cast<int>(nh);   // items per (batch, head)

  // Sequence numbers shared by all roles (each role counts them itself):
  //   it = item counter of this CTA;  g = K/V tile counter;  u = g * nh + h = score-tile step counter
  //   S / P' buffer = u & 1;  Q tile number = it * nh + h;  O accumulator number = it * nh + h
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA loader: Q and K (warp-converged, one
    // elected lane issues: see the PV issuer)
    {
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int qg = item % q_groups, bh = item / q_groups;
        for (uint32_t h = 0; h < nh; ++h) {
          const uint32_t qn = it * nh + h, qb = qn % nqb;
          mbar_wait(&bars[B_QEMPTY + qb], ((qn / nqb) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&bars[B_QFULL + qb], q_bytes);
            for (int c = 0; c < dchunks; ++c)
              tma_load_3d(s_q + qb * q_bytes + c * kChunkBytes, &tm_q, &bars[B_QFULL + qb], c * 64,
                          (qg * static_cast<int>(nh) + static_cast<int>(h)) * kTileQ, bh);
          }
          __syncwarp();
        }
        for (int j = 0; j < p.nkv; ++j) {
          for (int part = 0; part < kparts; ++part, ++g) {     // g counts ring entries here
            const uint32_t slot = g % nkb;
            mbar_wait(&bars[B_KEMPTY + slot], ((g / nkb) & 1) ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&bars[B_KFULL + slot], q_bytes);
              for (int c = 0; c < dchunks; ++c)
                tma_load_3d(s_k + slot * q_bytes + c * kChunkBytes, &tm_k, &bars[B_KFULL + slot], (part * dchunks + c) * 64,
                            j * kTileK, bh);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (PASS == 2 && warp == 2) {
    // ------------------------------------------------------------------ V loader (own ring, own warp)
    {
      uint32_t g = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int bh = item / q_groups;
        for (int j = 0; j < p.nkv; ++j, ++g) {
          const uint32_t slot = g % nvb;
          mbar_wait(&bars[B_VEMPTY + slot], ((g / nvb) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&bars[B_VFULL + slot], v_stage);
            for (int c = 0; c < 2; ++c)
              tma_load_3d(s_v + slot * v_stage + c * (p.dp * 128), &tm_v, &bars[B_VFULL + slot], j * kTileK + c * 64, 0, bh);
          }
          __syncwarp();
        }
      }
    }
  } else if (PASS == 2 && warp == 3) {
    // ------------------------------------------------------------------ PV issuer (own thread: every
    // mbarrier wait / commit costs the issuing thread ~100+ cycles of latency, so one thread issuing
    // both QK^T and PV serialises ~12 such operations per 128 x 128 step and becomes the critical path)
    // All 32 lanes run the loop in lock-step -- loop counters, barrier addresses and operand descriptors are then
    // warp-uniform and stay on the uniform datapath -- and ONE elected lane issues the tcgen05 instructions.  (As
    // `if (lane == 0) { loop }` the descriptors lived in per-thread registers and every MMA dragged a chain of
    // R2UR moves: ncu showed the QK^T issuer busy 75 % of the kernel and every other role waiting on it.)
    {
      const uint32_t idesc_o = umma_idesc_f16(kTileQ, p.dp);
      const uint64_t dp0 = umma_desc_sw128(smem_u32(s_p)), dv0 = umma_desc_sw128(smem_u32(s_v));
      const uint32_t cstep = kChunkBytes >> 4, vstep = v_stage >> 4, vcstep = static_cast<uint32_t>(p.dp * 128) >> 4;
      // O[on] += P'(u) V(g):  u = score-tile step, g = its K/V tile, on = O accumulator number,
      // first / last = first / last K tile of the item, lastq = last query half using V(g)
      auto issue_pv = [&](uint32_t u, uint32_t g, uint32_t on, bool first, bool last, bool lastq) {
        const uint32_t slot = g % nvb, pb = u & 1, ob = on % nob;
        mbar_wait(&bars[B_VFULL + slot], (g / nvb) & 1);
        if (p.ptm) {
          // P'(u) sits in tensor memory, written over the first 32 columns of each 64-column half of its own S tile:
          // A operand straight from TMEM (8 packed columns per K = 16 step).  With P' in shared memory every
          // M128 x N64 x K16 MMA read 6 KB of operands (48 cycles at the 128 B/clk shared-memory port, more than its
          // 32 cycles of math) and the softmax warps wrote another 32 KB per step through the same port: the port,
          // not the tensor pipe or the ALUs, bounded pass 2 (~170 KB per 128 x 128 step against 1885 cycles).
          const uint32_t sb = u % ns;
          mbar_wait(&bars[B_PFULL + sb], (u / ns) & 1);
          if (first) mbar_wait(&bars[B_OEMPTY + ob], ((on / nob) & 1) ^ 1);
          tc_fence_after();
          const uint64_t db0 = dv0 + slot * vstep;
          const uint32_t d_o = tmem_o + ob * p.dp, a_t = tmem_base + sb * kTileK;
          if (elect_one()) {
            DGQ_TR(3, u);                                             // PV of step u goes out
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                tc_mma_f16_ts(d_o, a_t + c * 64 + ks * 8, db0 + c * vcstep + 2 * ks, idesc_o, (!first || (c | ks) != 0) ? 1u : 0u);
            }
            if (lastq) tc_commit(&bars[B_VEMPTY + slot]);
            tc_commit(&bars[B_SEMPTY + sb + (TWO ? 3 * ((u & 1) ^ 1) : 0)]);   // the S / P' tile is free for the QK^T of step u + 3 (TWO: the other half's issuer)
            if (last) tc_commit(&bars[B_OFULL + ob]);
          }
          __syncwarp();
          return;
        }
        mbar_wait(&bars[B_PFULL + pb], (u >> 1) & 1);
        if (first) mbar_wait(&bars[B_OEMPTY + ob], ((on / nob) & 1) ^ 1);   // the epilogue drained this accumulator
        tc_fence_after();
        const uint64_t da0 = dp0 + pb * 2 * cstep, db0 = dv0 + slot * vstep;
        const uint32_t d_o = tmem_o + ob * p.dp;
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_f16(d_o, da0 + c * cstep + 2 * ks, db0 + c * vcstep + 2 * ks, idesc_o, (!first || (c | ks) != 0) ? 1u : 0u);
          }
          if (lastq) tc_commit(&bars[B_VEMPTY + slot]);
          tc_commit(&bars[B_PEMPTY + pb]);
          if (last) tc_commit(&bars[B_OFULL + ob]);
        }
        __syncwarp();
      };
      uint32_t u = 0, g = 0, it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        for (int j = 0; j < p.nkv; ++j, ++g) {
          for (uint32_t h = 0; h < nh; ++h, ++u)
            issue_pv(u, g, it * nh + h, j == 0, j == p.nkv - 1, h == nh - 1);
        }
      }
    }
  } else if (warp == 1 || (TWO && warp == Cfg::kQkWarpB)) {
    // ------------------------------------------------------------------ QK^T issuer (warp-converged, one elected
    // lane issues: see the PV issuer).  Ping-pong pass 2: warp 1 issues for query half 0, warp kQkWarpB for half 1.
    {
      constexpr bool two = TWO;
      const uint32_t h_first = (two && warp != 1) ? 1u : 0u, h_step = two ? 2u : 1u;
      const uint32_t idesc_s = umma_idesc_f16(kTileQ, kTileK);
      const uint64_t dq0 = umma_desc_sw128(smem_u32(s_q)), dk0 = umma_desc_sw128(smem_u32(s_k));
      const uint32_t cstep = kChunkBytes >> 4, qstep = q_bytes >> 4;      // descriptor address field: 16-byte units
      uint32_t u = 0, g = 0, it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        for (int j = 0; j < p.nkv; ++j, ++g, u += nh) {
          // S(u + h) = Q_h . K_hi (+ Q_h . K_lo): the parts arrive as separate ring entries
          if (nkb >= static_cast<uint32_t>(kparts)) {
            // both parts resident: one query half at a time, so S(u) is committed without waiting for the OTHER
            // half's S buffer -- the two softmax groups must stay decoupled (a part-outer order locks them in step)
            for (int part = 0; part < kparts; ++part) {
              const uint32_t gp = g * kparts + part;
              mbar_wait(&bars[B_KFULL + gp % nkb], (gp / nkb) & 1);
            }
            const uint32_t slot0 = (g * kparts) % nkb, slot1 = (g * kparts + kparts - 1) % nkb;
            for (uint32_t h = h_first; h < nh; h += h_step) {
              const uint32_t uu = u + h, qn = it * nh + h, qb = qn % nqb, sb = uu % ns;
              if (elect_one()) DGQ_TR(0, uu);                         // issuer reaches the step
              if (j == 0) mbar_wait(&bars[B_QFULL + qb], (qn / nqb) & 1);
              const uint32_t sset = (two && ns == 3) ? 3 * h : 0;     // barrier set of this query half
              if (two && ns == 3) {
                if (uu >= 3) mbar_wait(&bars[B_SEMPTY + sb + sset], ((uu - 3) / 6) & 1);   // step uu - 3 released the tile
              } else {
                mbar_wait(&bars[B_SEMPTY + sb], ((uu / ns) & 1) ^ 1);
              }
              tc_fence_after();
              const uint64_t da0 = dq0 + qb * qstep, db0 = dk0 + slot0 * qstep, db1 = dk0 + slot1 * qstep;
              const uint32_t d_s = tmem_base + sb * kTileK;
              if (elect_one()) {
                DGQ_TR(1, uu);                                        // waits done: first MMA goes out
                for (int c = 0; c < dchunks; ++c) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    tc_mma_f16(d_s, da0 + c * cstep + 2 * ks, db0 + c * cstep + 2 * ks, idesc_s, (c | ks) != 0 ? 1u : 0u);
                }
                if (kparts == 2) {
                  for (int c = 0; c < dchunks; ++c) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                      tc_mma_f16(d_s, da0 + c * cstep + 2 * ks, db1 + c * cstep + 2 * ks, idesc_s, 1u);
                  }
                }
                tc_commit(&bars[B_SFULL + sb + sset]);
                DGQ_TR(2, uu);                                        // S-full commit issued
                if (j == p.nkv - 1) tc_commit(&bars[B_QEMPTY + qb]);         // last read of this Q tile
                if (h == nh - 1 || two) {                                    // (this issuer's) last read of this K tile
                  tc_commit(&bars[B_KEMPTY + slot0]);
                  if (kparts == 2) tc_commit(&bars[B_KEMPTY + slot1]);
                }
              }
              __syncwarp();
            }
            continue;
          }
          // hi and lo share ONE buffer (dp = 128 / 192 with the split): part by part
          for (int part = 0; part < kparts; ++part) {
            const uint32_t gp = g * kparts + part, slot = gp % nkb;
            mbar_wait(&bars[B_KFULL + slot], (gp / nkb) & 1);
            for (uint32_t h = h_first; h < nh; h += h_step) {
              const uint32_t uu = u + h, qn = it * nh + h, qb = qn % nqb, sb = uu % ns;
              if (part == 0) {
                if (j == 0) mbar_wait(&bars[B_QFULL + qb], (qn / nqb) & 1);
                mbar_wait(&bars[B_SEMPTY + sb], ((uu / ns) & 1) ^ 1);
              }
              tc_fence_after();
              const uint64_t da0 = dq0 + qb * qstep, db0 = dk0 + slot * qstep;
              const uint32_t d_s = tmem_base + sb * kTileK;
              if (elect_one()) {
                for (int c = 0; c < dchunks; ++c) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    tc_mma_f16(d_s, da0 + c * cstep + 2 * ks, db0 + c * cstep + 2 * ks, idesc_s, (part | c | ks) != 0 ? 1u : 0u);
                }
                if (part == kparts - 1) {
                  tc_commit(&bars[B_SFULL + sb]);
                  if (j == p.nkv - 1) tc_commit(&bars[B_QEMPTY + qb]);       // last read of this Q tile
                }
                if (h == nh - 1 || two) tc_commit(&bars[B_KEMPTY + slot]);   // (this issuer's) last read of this ring entry
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    // warp -> TMEM lane quarter (warp & 3) and column split: one thread per (row, 128 / kSplit columns)
    // of each 128 x 128 score tile
    const int quad = warp & 3;
    const int grp = PP ? (warp - Cfg::kFirstSoftmaxWarp) >> 3 : 0;          // PP: the query half this warp owns
    const int half = ((warp - Cfg::kFirstSoftmaxWarp) >> 2) & (kSplit - 1);   // column split index, 0 .. kSplit - 1
    const int col_lo = half * (128 / kSplit);     // first score column of this thread inside a tile
    const int etid = (threadIdx.x - 32 * Cfg::kFirstSoftmaxWarp) & (kGroupThreads - 1);   // index inside the group
    const int row = quad * 32 + lane;             // row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const bool partial_last = (p.s % kTileK) != 0;
    const uint32_t sp_base = smem_u32(s_p);
    uint32_t u = 0, it = 0;

    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int qg = item % q_groups, bh = item / q_groups;
      int tq[2];
      bool row_ok[2];
      size_t ridx[2];
      float al[2];                                  // alpha of the row: scale * log2(e) * (delta of its Q quantizer)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        tq[h] = (qg * static_cast<int>(nh) + h) * kTileQ + row;   // query index
        row_ok[h] = h < static_cast<int>(nh) && tq[h] < p.t;
        ridx[h] = static_cast<size_t>(bh) * p.t + tq[h];
        al[h] = p.alpha * ((p.q_scale != nullptr && row_ok[h]) ? __ldg(p.q_scale + tq[h] % p.q_period) : 1.0f);
      }

      // O accumulator `ob`, columns [c_lo, c_hi) of this thread's row: O * out_scale (+ p0 * v0), the to_out quantizer, store
      const int head = bh % p.heads, bb = bh / p.heads;
      // `stage` != nullptr (ping-pong path): results go through a 4 KB per-warp, XOR-swizzled shared-memory tile
      // (carved from the group's idle P' buffer) and leave as 128-byte (fp32) / 64-byte (fp16) row segments;
      // thread-per-row stores of 16 bytes cost 32 L1 wavefronts per instruction and made this epilogue 40-45 % of
      // the cross-attention kernels.
      auto store_o = [&](int tqh, bool rok, float pp0, const float* v0, float oscale, uint32_t ob, int c_lo, int c_hi,
                         uint8_t* stage) {
          const size_t orow_idx = static_cast<size_t>(bb) * p.t + tqh;
          const size_t ooff = orow_idx * p.ldo + head * p.d;
          __half* orow = static_cast<__half*>(p.out) + ooff;
          float* orow32 = static_cast<float*>(p.out) + ooff;
          uint8_t* orow8 = static_cast<uint8_t*>(p.out) + ooff;
          const bool out_u8 = p.oq_emit_int == 2;   // u8 codes for a kind::i8 to_out GEMM (ldo in bytes)
          float rd = 1.f, rz = 0.f, ri = 1.f;       // row-indexed / scalar output quantizer
          if (rok && (p.oq_mode == DGQ_Q_SCALAR || p.oq_mode == DGQ_Q_ROWWISE)) {
            const int jq = p.oq_mode == DGQ_Q_ROWWISE ? static_cast<int>(orow_idx % p.oq_period) : 0;
            rd = __ldg(p.oq_delta + jq); rz = __ldg(p.oq_zp + jq);
            ri = p.oq_inv != nullptr ? __ldg(p.oq_inv + jq) : rcp_rn_slow(rd);
          }
          const int rows_ok = p.t - (tqh - lane);   // rows of this warp's 32 that exist (tqh - lane = its first query row)
          for (int c = c_lo; c < c_hi; c += 16) {
            uint32_t r[16];
            tmem_ld_32x16(tmem_o + ob * p.dp + lane_addr + c, r);
            tc_wait_ld();
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const int d0 = c + v * 8;
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = 0.f;
              if (rok && d0 < p.d) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  f[i] = __uint_as_float(r[v * 8 + i]) * oscale;
                  if (p.start_peak) f[i] = fmaf(pp0, v0[d0 + i], f[i]);
                }
                if (p.oq_mode == DGQ_Q_KWISE) {
                  const int k0 = head * p.d + d0;
                  float qd[8], qz[8], qi[8], lo[8], hi[8];
                  ldg8(p.oq_delta + k0, qd);
                  ldg8(p.oq_zp + k0, qz);
                  if (p.oq_inv != nullptr) {
                    ldg8(p.oq_inv + k0, qi);
                  } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) qi[i] = rcp_rn_slow(qd[i]);
                  }
                  uaq_bounds<8>(qz, p.oq_qmax, lo, hi);
                  uaq_lean_lh<false, 8>(f, qd, qi, lo, hi);
                } else if (p.oq_mode != DGQ_Q_NONE) {
                  if (p.oq_emit_int) uaq_lean1_lh<true, 8>(f, rd, ri, -rz, __fsub_rn(p.oq_qmax, rz));
                  else uaq_lean1_lh<false, 8>(f, rd, ri, -rz, __fsub_rn(p.oq_qmax, rz));
                  if (out_u8) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] += rz;          // the u8 code itself, exact in fp16
                  }
                }
              }
              if (stage == nullptr) {
                if (rok && d0 < p.d) {
                  if (p.out_is_f32) {
                    *reinterpret_cast<float4*>(orow32 + d0) = make_float4(f[0], f[1], f[2], f[3]);
                    *reinterpret_cast<float4*>(orow32 + d0 + 4) = make_float4(f[4], f[5], f[6], f[7]);
                  } else if (out_u8) {
                    const uint4 h = pack8(f);
                    *reinterpret_cast<uint2*>(orow8 + d0) = make_uint2(halves4_to_u8(h.x, h.y), halves4_to_u8(h.z, h.w));
                  } else {
                    *reinterpret_cast<uint4*>(orow + d0) = pack8(f);
                  }
                }
              } else {
                const int j8 = ((c - c_lo) & 16) / 8 + v;            // 8-column group inside the 32-column block
                if (p.out_is_f32) {
                  *reinterpret_cast<float4*>(stage + lane * 128 + (((2 * j8) ^ (lane & 7)) << 4)) = make_float4(f[0], f[1], f[2], f[3]);
                  *reinterpret_cast<float4*>(stage + lane * 128 + (((2 * j8 + 1) ^ (lane & 7)) << 4)) = make_float4(f[4], f[5], f[6], f[7]);
                } else {
                  *reinterpret_cast<uint4*>(stage + lane * 64 + ((j8 ^ ((lane >> 1) & 3)) << 4)) = pack8(f);
                }
              }
            }
            if (stage != nullptr && (((c - c_lo) & 16) != 0 || c + 16 >= c_hi)) {
              // a 32-column block (or the tail) is staged: write it out as row segments
              const int cb = c_lo + ((c - c_lo) & ~31);              // first column of the block
              __syncwarp();
              // element offset of (row 0 of this warp, column cb) from this thread's own row pointer
              const ptrdiff_t row0 = -static_cast<ptrdiff_t>(lane) * p.ldo;
              if (p.out_is_f32) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int rr = i * 4 + (lane >> 3), ch = lane & 7;
                  const float4 x = *reinterpret_cast<const float4*>(stage + rr * 128 + ((ch ^ (rr & 7)) << 4));
                  const int col = cb + ch * 4;
                  if (rr < rows_ok && col < c_hi && col < p.d)
                    *reinterpret_cast<float4*>(orow32 + row0 + static_cast<ptrdiff_t>(rr) * p.ldo + col) = x;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int rr = i * 8 + (lane >> 2), ch = lane & 3;
                  const uint4 x = *reinterpret_cast<const uint4*>(stage + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
                  const int col = cb + ch * 8;
                  if (rr < rows_ok && col < c_hi && col < p.d) {
                    if (out_u8)
                      *reinterpret_cast<uint2*>(orow8 + row0 + static_cast<ptrdiff_t>(rr) * p.ldo + col) =
                          make_uint2(halves4_to_u8(x.x, x.y), halves4_to_u8(x.z, x.w));
                    else
                      *reinterpret_cast<uint4*>(orow + row0 + static_cast<ptrdiff_t>(rr) * p.ldo + col) = x;
                  }
                }
              }
              __syncwarp();
            }
          }
      };
      if constexpr (PASS == 2 && PP) {
        // ---------------------------------------------------------------- pass 2, ping-pong groups (nh == 2):
        // this warp's group owns query half `grp` of the item: S / P' buffer grp, O accumulator grp.
        // u counts the K/V tiles of this CTA (= the phase of the group's buffers).
        const int tqh = (qg * 2 + grp) * kTileQ + row;
        const bool rok = tqh < p.t;
        const size_t rix = static_cast<size_t>(bh) * p.t + tqh;
        const float alpha = p.alpha * ((p.q_scale != nullptr && rok) ? __ldg(p.q_scale + tqh % p.q_period) : 1.0f);
        float delta = 1.0f;
        if (MODE != DGQ_MAP_NONE) delta = p.real_time ? p.gmax[0] : __ldg(p.delta);
        const float lg_delta = MODE != DGQ_MAP_NONE ? log2f(delta) : 0.f;
        const float qcap = 64.f;                  // power of two: see map_chunk
        const float beta = rok ? (p.row_max[rix] + log2f(p.row_sum[rix])) : 0.f;
        const float gamma = beta + lg_delta;
        float p0 = 0.f;                           // un-quantised start-peak probability of this row
        float* v0 = s_v0 + (grp * 2 + (it & 1)) * 192;
        if (p.start_peak) {
          for (int dd = etid; dd < p.dp; dd += kGroupThreads)
            v0[dd] = __half2float(p.vt[(static_cast<size_t>(bh) * p.dp + dd) * p.sp]);
        }
        const uint32_t pbuf = grp;                // the group's P' buffer and O accumulator
        const int s_len = (CODES && !rok) ? 0 : p.s;
        uint8_t* code_row = CODES ? p.codes + rix * p.s : nullptr;
        // this thread's 64 score columns are one [128 x 64] SW128 sub-tile of the P' buffer
        const uint32_t sub = sp_base + pbuf * 2 * kChunkBytes + half * kChunkBytes;
        for (int j = 0; j < p.nkv; ++j, ++u) {
          const uint32_t ph = u & 1;              // phase of the group's P' buffer
          const uint32_t uu = 2 * u + grp, sb = uu % ns;   // CTA-wide step number -> its S buffer
          if (etid == 0) DGQ_TR(4, uu);                               // group asks for S(uu)
          if (TWO && ns == 3) mbar_wait(&bars[B_SFULL + sb + 3 * grp], (uu / 6) & 1);   // this half's own barrier set
          else mbar_wait(&bars[B_SFULL + sb], (uu / ns) & 1);
          if (etid == 0) DGQ_TR(5, uu);                               // ... has it
          tc_fence_after();
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + lane_addr + sb * kTileK + col_lo + cc * 32, r);
            tc_wait_ld();
            if (cc == 1 && !p.ptm) {              // all of this thread's scores are in registers: free the S tile
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bars[B_SEMPTY + sb + ((TWO && ns == 3) ? 3 * (grp ^ 1) : 0)]);   // to the next user's issuer
            }
            uint32_t h2[16];
            map_chunk<MODE, CODES>(r, h2, alpha, gamma, qcap, p.qmax, j * kTileK + col_lo + cc * 32, s_len, code_row);
            if (cc == 0) {
              if (p.start_peak && j == 0 && half == 0) {
                p0 = ex2_approx(fmaf(alpha, __uint_as_float(r[0]), -beta));
                h2[0] &= 0xFFFF0000u;             // column 0 leaves the MMA; added back in the epilogue
              }
              if (!p.ptm) mbar_wait(&bars[B_PEMPTY + pbuf], ph ^ 1);   // the PV of the previous K tile has consumed this P' buffer
            }
            if (p.ptm) {
              // 32 fp16 = 16 packed words over the scores they came from: columns [col_lo + 16 cc, + 16) of the row
              // (this thread's own, already consumed, quarter of the S tile)
              tmem_st_32x16(tmem_base + lane_addr + sb * kTileK + col_lo + cc * 16, h2);
              continue;
            }
#pragma unroll
            for (int v = 0; v < 4; ++v)
              st_shared_v4(sub + sw128_offset(row, cc * 4 + v), h2[4 * v], h2[4 * v + 1], h2[4 * v + 2], h2[4 * v + 3]);
          }
          if (p.ptm) {
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_PFULL + sb]);
            if (etid == 0) DGQ_TR(6, uu);                             // P'(uu) published
            continue;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[B_PFULL + pbuf]);
        }
        // ---- epilogue of this group's half; the other group keeps the tensor pipe busy meanwhile
        float pp0 = p0;
        if (p.start_peak) {
          if (half == 0) s_x[grp * 128 + row] = pp0;
          softmax_bar_sync<kGroupThreads>(1 + grp);
          pp0 = s_x[grp * 128 + row];
        }
        const uint32_t ob = grp;                  // O accumulator number it * 2 + grp, two accumulators
        mbar_wait(&bars[B_OFULL + ob], it & 1);
        tc_fence_after();
        const int dsplit = p.dp >> 1;
        // staging tile: this warp's 4 KB of the group's P' buffer (every PV of the item has retired: OFULL)
        uint8_t* stage = s_p + grp * 2 * kChunkBytes + ((warp - Cfg::kFirstSoftmaxWarp) & 7) * 4096;
        store_o(tqh, rok, pp0, v0, MODE == DGQ_MAP_NONE ? 1.0f : delta, ob, half * dsplit, (half + 1) * dsplit, stage);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_OEMPTY + ob]);
        if (p.start_peak) softmax_bar_sync<kGroupThreads>(1 + grp);   // s_x / v0 are rewritten next
        continue;
      }
      if (PASS == 1) {
        float M[2] = {-INFINITY, -INFINITY}, Mx[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
        for (int j = 0; j < p.nkv; ++j) {
          const bool mask = partial_last && j == p.nkv - 1;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h >= static_cast<int>(nh)) break;
            const uint32_t sb = u & 1;
            mbar_wait(&bars[B_SFULL + sb], (u >> 1) & 1);
            ++u;
            tc_fence_after();
            uint32_t r[kCh][32];
#pragma unroll
            for (int cc = 0; cc < kCh; ++cc) tmem_ld_32x32(tmem_base + lane_addr + sb * kTileK + col_lo + cc * 32, r[cc]);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_SEMPTY + sb]);   // scores are in registers: free the S tile
#pragma unroll
            for (int cc = 0; cc < kCh; ++cc) {
              const int c = half * kCh + cc;
              // row max on the raw scores (alpha > 0), then ex2(s * alpha - M) as ONE FMA + MUFU per element
              float x[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[cc][i]);
              if (mask) {
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = (j * kTileK + c * 32 + i < p.s) ? x[i] : -INFINITY;
              }
              float cm = x[1];
#pragma unroll
              for (int i = 2; i < 32; ++i) cm = fmaxf(cm, x[i]);
              const float cmx = cm * al[h];         // excludes element 0 of this chunk
              cm = fmaxf(cm, x[0]) * al[h];
              Mx[h] = fmaxf(Mx[h], (j == 0 && c == 0) ? cmx : cm);
              const float Mn = fmaxf(M[h], cm);
              if (Mn > -INFINITY) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float e = fmaf(x[i], al[h], -Mn);
                  acc += (i % kPolyEvery) == kPolyEvery - 1 ? ex2_poly(e) : ex2_approx(e);
                }
                l[h] = l[h] * ex2_approx(M[h] - Mn) + acc;
                M[h] = Mn;
              }
            }
          }
        }
        // combine the column splits of each row
        static_assert(PASS != 1 || kSplit == 2, "pass 1 exchanges one partial per row");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h >= static_cast<int>(nh)) break;
          if (half == 1) { s_x[row] = M[h]; s_x[128 + row] = l[h]; s_x[256 + row] = Mx[h]; }
          softmax_bar_sync<kSoftmaxThreads>();
          if (half == 0) {
            const float M1 = s_x[row], l1 = s_x[128 + row], Mx1 = s_x[256 + row];
            const float Mn = fmaxf(M[h], M1);
            float ll = l[h];
            if (Mn > -INFINITY) ll = ll * ex2_approx(M[h] - Mn) + l1 * ex2_approx(M1 - Mn);
            const float Mxx = fmaxf(Mx[h], Mx1);
            float rp = 0.f;
            if (row_ok[h]) {
              p.row_max[ridx[h]] = Mn;
              p.row_sum[ridx[h]] = ll;
              rp = (p.start_peak ? ex2_approx(Mxx - Mn) : 1.0f) / ll;
            }
            for (int o = 16; o > 0; o >>= 1) rp = fmaxf(rp, __shfl_xor_sync(0xffffffffu, rp, o));
            if (lane == 0 && p.real_time) atomicMax(reinterpret_cast<int*>(p.gmax), __float_as_int(rp));
          }
          softmax_bar_sync<kSoftmaxThreads>();      // s_x is rewritten next
        }
      } else {
        // ---------------------------------------------------------------- pass 2
        float delta = 1.0f;
        if (MODE != DGQ_MAP_NONE) delta = p.real_time ? p.gmax[0] : __ldg(p.delta);
        const float lg_delta = MODE != DGQ_MAP_NONE ? log2f(delta) : 0.f;
        const float qcap = 64.f;                  // power of two: see map_chunk
        float beta[2], p0[2] = {0.f, 0.f};        // p0: un-quantised start-peak probability of this row
#pragma unroll
        for (int h = 0; h < 2; ++h) beta[h] = row_ok[h] ? (p.row_max[ridx[h]] + log2f(p.row_sum[ridx[h]])) : 0.f;
        float* v0 = s_v0 + (it & 1) * 192;
        if (p.start_peak) {
          for (int dd = etid; dd < p.dp; dd += kSoftmaxThreads)
            v0[dd] = __half2float(p.vt[(static_cast<size_t>(bh) * p.dp + dd) * p.sp]);
        }
        for (int j = 0; j < p.nkv; ++j) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h >= static_cast<int>(nh)) break;
            const uint32_t sb = u & 1, ph = (u >> 1) & 1;
            ++u;
            mbar_wait(&bars[B_SFULL + sb], ph);
            tc_fence_after();
            uint32_t r[kCh][32];
#pragma unroll
            for (int cc = 0; cc < kCh; ++cc) tmem_ld_32x32(tmem_base + lane_addr + sb * kTileK + col_lo + cc * 32, r[cc]);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_SEMPTY + sb]);   // scores are in registers: free the S tile
            const int s_len = (CODES && !row_ok[h]) ? 0 : p.s;
            uint8_t* code_row = CODES ? p.codes + ridx[h] * p.s : nullptr;
            const float gamma = beta[h] + lg_delta;
            uint32_t h2[kCh][16];
#pragma unroll
            for (int cc = 0; cc < kCh; ++cc)
              map_chunk<MODE, CODES>(r[cc], h2[cc], al[h], gamma, qcap, p.qmax, j * kTileK + col_lo + cc * 32, s_len,
                                     code_row);
            if (p.start_peak && j == 0 && half == 0) {
              p0[h] = ex2_approx(fmaf(al[h], __uint_as_float(r[0][0]), -beta[h]));
              h2[0][0] &= 0xFFFF0000u;              // column 0 leaves the MMA; added back in the epilogue
            }
            mbar_wait(&bars[B_PEMPTY + sb], ph ^ 1);   // PV of step u - 2 has consumed this P' buffer
            // P' tile = two [128 x 64] SW128 sub-tiles; score column col -> sub-tile col / 64, 16-byte chunk (col % 64) / 8
            const uint32_t sub = sp_base + sb * 2 * kChunkBytes + (col_lo >> 6) * kChunkBytes;
            const int ch0 = (col_lo & 63) >> 3;
#pragma unroll
            for (int cc = 0; cc < kCh; ++cc) {
#pragma unroll
              for (int v = 0; v < 4; ++v)
                st_shared_v4(sub + sw128_offset(row, ch0 + cc * 4 + v), h2[cc][4 * v], h2[cc][4 * v + 1], h2[cc][4 * v + 2],
                             h2[cc][4 * v + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_PFULL + sb]);
          }
        }
        // ---- epilogue: O * out_scale (+ p0 * v0) -> out; the column splits share the dp columns
        const float oscale = MODE == DGQ_MAP_NONE ? 1.0f : delta;
        const int dsplit = p.dp / kSplit;         // 16-column multiples: dp = 64 / 128 / 192, kSplit = 4
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h >= static_cast<int>(nh)) break;
          float pp0 = p0[h];
          if (p.start_peak) {
            if (half == 0) s_x[row] = pp0;
            softmax_bar_sync<kSoftmaxThreads>();
            pp0 = s_x[row];
          }
          const uint32_t on = it * nh + h, ob = on % nob;
          mbar_wait(&bars[B_OFULL + ob], (on / nob) & 1);
          tc_fence_after();
          store_o(tq[h], row_ok[h], pp0, v0, oscale, ob, half * dsplit, (half + 1) * dsplit, nullptr);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[B_OEMPTY + ob]);
          if (p.start_peak) softmax_bar_sync<kSoftmaxThreads>();   // s_x is rewritten next
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
  DGQ_CHECK_ARG(a->q_scale == nullptr || a->q_scale_period > 0);
  DGQ_CHECK_ARG(a->out_q.emit_int != 2 || (a->out_q.mode != DGQ_Q_NONE && !a->out_is_f32 && a->ldo % 16 == 0));

  AttnDev p;
  p.b = a->b; p.heads = a->heads; p.t = a->t; p.s = a->s; p.d = a->d; p.dp = a->dp;
  p.nkv = (a->s + kTileK - 1) / kTileK;
  // Work item = nh 128-row query halves of one (batch, head): with nh = 2 every K/V tile is loaded once
  // for both halves (the kernel is L2 -> smem bound otherwise).  TMEM (512 columns): S 2 x 128 + O 2 x dp,
  // so dp = 192 runs nh = 1 with a single O accumulator.  smem (227 KB), in 16 KB units for dp = 64:
  // Q 4 + K 3 + V 2 + P' 4; dp = 128: Q 2 + K 1 + V 1 (32 KB units) + P'; dp = 192: Q 1 + K 1 + V 1 (48 KB) + P'.
  p.q_tiles = (a->t + kTileQ - 1) / kTileQ;
  p.k_split = a->k_split ? 1 : 0;
  p.q_scale = a->q_scale; p.q_period = a->q_scale_period > 0 ? a->q_scale_period : 1;
  p.nh = (a->dp <= 128 && p.q_tiles > 1) ? 2 : 1;
  if (!p.k_split) {
    p.nq_buf = a->dp <= 64 ? 4 : (a->dp <= 128 ? 2 : 1);
    p.nk_buf = a->dp <= 64 ? 3 : 1;
  } else {     // K tiles travel as two ring entries (hi, lo) of one Q-tile size each; dp > 64: hi and lo share one buffer
    p.nq_buf = a->dp <= 128 ? 2 : 1;
    p.nk_buf = a->dp <= 64 ? 4 : 1;
  }
  p.nv_buf = a->dp <= 64 ? 2 : 1;
  if (p.k_split && a->dp <= 64 && p.nkv == 1) {   // cross-attention: one K/V tile per item, latency-bound -- spend
    p.nq_buf = 4;                                   // the shared memory on the NEXT item's Q tiles (16 KB units:
    p.nv_buf = 1;                                   // Q 4 + K 4 + V 1 + P' 4 = 208 KB)
  }
  p.no_buf = a->dp <= 128 ? 2 : 1;
  p.ns = (p.nh == 2 && a->dp <= 64) ? 3 : 2;
  static int ptm_on = -1;       // DGQ_ATTN_PTMEM=0: P' through shared memory everywhere (A/B runs)
  if (ptm_on < 0) {
    const char* env = getenv("DGQ_ATTN_PTMEM");
    ptm_on = (env == nullptr || atoi(env) != 0) ? 1 : 0;
  }
  p.ptm = (ptm_on && p.ns == 3) ? 1 : 0;
  p.items = a->b * a->heads * ((p.q_tiles + p.nh - 1) / p.nh);
  p.alpha = a->scale * 1.4426950408889634f;
  p.map_mode = a->map_mode; p.real_time = a->real_time; p.start_peak = a->start_peak;
  p.delta = a->delta; p.qmax = a->qmax;
  p.row_max = a->row_max; p.row_sum = a->row_sum; p.gmax = a->gmax;
  p.vt = static_cast<const __half*>(a->vt); p.sp = a->sp;
  p.out = a->out; p.ldo = a->ldo; p.out_is_f32 = a->out_is_f32; p.codes = a->codes;
  p.oq_delta = a->out_q.delta; p.oq_zp = a->out_q.zp; p.oq_inv = a->out_q.inv_delta; p.oq_mode = a->out_q.mode;
  p.oq_period = a->out_q.period > 0 ? a->out_q.period : 1; p.oq_emit_int = a->out_q.emit_int;
  p.oq_qmax = a->out_q.qmax;

  const uint64_t bh = static_cast<uint64_t>(a->b) * a->heads;
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_3d(&tq, a->q, bh, a->t, a->dp, kTileQ);
  if (rc != 0) return rc;
  rc = make_tmap_3d(&tk, a->k, bh, a->s, p.k_split ? 2 * a->dp : a->dp, kTileK);   // k_split: [.., s, hi dp | lo dp]
  if (rc != 0) return rc;
  rc = make_tmap_3d(&tv, a->vt, bh, a->dp, a->sp, a->dp);
  if (rc != 0) return rc;

  const uint32_t q_bytes = (a->dp / 64) * kChunkBytes;
  const uint32_t tail = 1024 + B_COUNT * 8 + 16 + 4 * 192 * 4 + 3 * 128 * 4 + 64;
  AttnDev p1 = p;                       // pass 1 keeps two CTAs per SM: a smaller Q ring, 256 TMEM columns
  if (p1.nq_buf > 2) p1.nq_buf = 2;
  p1.ns = 2;
  p1.ptm = 0;
  const uint32_t smem1 = q_bytes * (p1.nq_buf + p1.nk_buf) + tail;
  const uint32_t smem2 = q_bytes * (p.nq_buf + p.nk_buf) + p.nv_buf * 2 * a->dp * 128 + 4 * kChunkBytes + tail;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnDev);
  KernelFn k1 = attention_kernel<1, 0, false, false>;
  KernelFn k2;
  const bool cd = a->codes != nullptr;
  const bool pp = p.nh == 2;            // two query halves per item: one softmax warp group per half
  // a second QK^T issuer for long self-attention with the log2 map (DGQ_ATTN_TWO=0 switches it off: A/B runs)
  static int two_on = -1;
  if (two_on < 0) {
    const char* env = getenv("DGQ_ATTN_TWO");
    two_on = (env == nullptr || atoi(env) != 0) ? 1 : 0;
  }
  const bool two = two_on && pp && !cd && a->map_mode == DGQ_MAP_LOG2 && p.nkv >= 4;
  switch (a->map_mode) {
    case DGQ_MAP_LOG2:
      k2 = cd ? (pp ? attention_kernel<2, DGQ_MAP_LOG2, true, true> : attention_kernel<2, DGQ_MAP_LOG2, true, false>)
              : (pp ? (two ? attention_kernel<2, DGQ_MAP_LOG2, false, true, true> : attention_kernel<2, DGQ_MAP_LOG2, false, true>)
                    : attention_kernel<2, DGQ_MAP_LOG2, false, false>);
      break;
    case DGQ_MAP_UNIFORM:
      k2 = cd ? (pp ? attention_kernel<2, DGQ_MAP_UNIFORM, true, true> : attention_kernel<2, DGQ_MAP_UNIFORM, true, false>)
              : (pp ? attention_kernel<2, DGQ_MAP_UNIFORM, false, true> : attention_kernel<2, DGQ_MAP_UNIFORM, false, false>);
      break;
    default:
      k2 = pp ? attention_kernel<2, DGQ_MAP_NONE, false, true> : attention_kernel<2, DGQ_MAP_NONE, false, false>;
      break;
  }
  // every instantiation gets the maximum it can ever need once (227 KB opt-in)
  static PerDeviceOnce attr;
  int dev;
  if (!attr.done(&dev)) {
    KernelFn all[] = {attention_kernel<1, 0, false, false>,
                      attention_kernel<2, DGQ_MAP_LOG2, true, false>, attention_kernel<2, DGQ_MAP_LOG2, true, true>,
                      attention_kernel<2, DGQ_MAP_LOG2, false, false>, attention_kernel<2, DGQ_MAP_LOG2, false, true>,
                      attention_kernel<2, DGQ_MAP_LOG2, false, true, true>,
                      attention_kernel<2, DGQ_MAP_UNIFORM, true, false>, attention_kernel<2, DGQ_MAP_UNIFORM, true, true>,
                      attention_kernel<2, DGQ_MAP_UNIFORM, false, false>, attention_kernel<2, DGQ_MAP_UNIFORM, false, true>,
                      attention_kernel<2, DGQ_MAP_NONE, false, false>, attention_kernel<2, DGQ_MAP_NONE, false, true>};
    for (KernelFn f : all) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    attr.mark(dev);
  }
  if (smem2 > 232448) return DGQ_ERR_INVALID_VALUE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // persistent CTAs: pass 1 fits two per SM (S only in TMEM), pass 2 one
  const int grid1 = p.items < 2 * kNumSMs ? p.items : 2 * kNumSMs;
  const int grid2 = p.items < kNumSMs ? p.items : kNumSMs;
  if (a->real_time) {
    cudaError_t e = cudaMemsetAsync(a->gmax, 0, sizeof(float), s);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  k1<<<grid1, AttCfg<1>::kThreads, smem1, s>>>(tq, tk, tv, p1);
  k2<<<grid2, two ? AttCfg<2, true>::kThreads : AttCfg<2>::kThreads, smem2, s>>>(tq, tk, tv, p);
  DGQ_RETURN_LAST_ERROR();
}

#ifdef DGQ_ATTN_TRACE
extern "C" int dgq_attn_trace_dump(long long* host) {
  return static_cast<int>(cudaMemcpyFromSymbol(host, dgq::g_trace, sizeof(dgq::g_trace)));
}
#endif
