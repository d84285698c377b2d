// Micro-probes that settle design questions on the B200 itself (results: profiles/r2_probes.txt).
//   mma   : issue rate of tcgen05.mma kind::f16 vs kind::i8 (operands resident in smem, no loads): the tensor-pipe
//           ceiling each kind can reach, per tile shape -- the "i8 peak measured the same way" of SURVEY.md 8d
//   ldtm  : tcgen05.ld throughput per SM for 4 / 8 / 16 warps (bounds the attention softmax passes and every epilogue)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_bin/probes scripts/probes.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../dgq_b200/csrc/ptx.cuh"

using namespace dgq;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e__ = (x);                                                                 \
    if (e__ != cudaSuccess) {                                                              \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);     \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

// KIND 0: f16 (K = 16 halves per instruction), 1: i8 (K = 32 bytes per instruction).  CTAS: cta_group.
template <int KIND, int CTAS>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters, int n, unsigned long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_a = smem;                 // [128 rows][128 B]
  uint8_t* s_b = smem + 16384;         // [256 rows][128 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_pair(slot, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(slot, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = KIND == 0 ? umma_idesc_f16(128 * CTAS, n) : umma_idesc_i8(128 * CTAS, n, false, true);
    const uint64_t da = umma_desc_sw128(smem_u32(s_a)), db = umma_desc_sw128(smem_u32(s_b));
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (it & 1) * 256;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (KIND == 0) {
          if (CTAS == 2) tc_mma_f16_pair(d, da + 2 * ks, db + 2 * ks, idesc, 1u);
          else tc_mma_f16(d, da + 2 * ks, db + 2 * ks, idesc, 1u);
        } else {
          if (CTAS == 2) tc_mma_i8_pair(d, da + 2 * ks, db + 2 * ks, idesc, 1u);
          else tc_mma_i8(d, da + 2 * ks, db + 2 * ks, idesc, 1u);
        }
      }
    }
    if (CTAS == 2) tc_commit_pair(&bar[0]); else tc_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    const unsigned long long t1 = clock64();
    if (blockIdx.x == 0) cyc[0] = t1 - t0;
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (CTAS == 2) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

template <int KIND, int CTAS>
static void run_mma(int n, int iters, unsigned long long* d_cyc) {
  const int smem = 16384 + 32768 + 1024 + 64;
  CK(cudaFuncSetAttribute(mma_peak_kernel<KIND, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148, 1, 1);
  cfg.blockDim = dim3(128, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  unsigned long long cyc = 0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, mma_peak_kernel<KIND, CTAS>, iters, n, d_cyc));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) { best = ms; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); }
  }
  const int kper = KIND == 0 ? 16 : 32;
  const double issuers = 148.0 / CTAS;
  const double macs = issuers * iters * 4.0 * (128.0 * CTAS) * n * kper;
  printf("mma kind::%s cta_group::%d M=%d N=%3d : %8.3f ms  %7.1f T%s/s  %6.1f cycles per instruction (SM clock)\n",
         KIND == 0 ? "f16" : "i8 ", CTAS, 128 * CTAS, n, best, 2.0 * macs / (best * 1e-3) / 1e12, KIND == 0 ? "FLOP" : "OP",
         static_cast<double>(cyc) / (iters * 4.0));
}

// every warp reads its own lane quarter, 32 columns per tcgen05.ld; `depth` loads in flight before the wait
template <int DEPTH>
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, unsigned long long* cyc, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) * 64) % 512;
  uint32_t acc = 0;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[DEPTH][32];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) tmem_ld_32x32(base + ((it * DEPTH + d) & 1) * 32, r[d]);
    tc_wait_ld();
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) acc ^= r[d][0] ^ r[d][13] ^ r[d][31];
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// tcgen05.ld while the tensor core is accumulating into the OTHER half of tensor memory: warp 0 = MMA issuer
// (kind::f16 or kind::i8, M = 128, N = 256, columns 0..255), warps 1.. read columns 256..511 in a loop -- what a
// double-buffered GEMM epilogue does.  Reports both rates.
template <int KIND>
__global__ void __launch_bounds__(544, 1) ldtm_mma_kernel(int mma_iters, int ld_iters, unsigned long long* cyc, uint32_t* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  uint32_t acc = 0;
  if (warp == 0) {
    if (threadIdx.x == 0) {
      const uint32_t idesc = KIND == 0 ? umma_idesc_f16(128, 256) : umma_idesc_i8(128, 256, false, true);
      const uint64_t da = umma_desc_sw128(smem_u32(s_a)), db = umma_desc_sw128(smem_u32(s_b));
      const unsigned long long t0 = clock64();
      for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (KIND == 0) tc_mma_f16(tmem, da + 2 * ks, db + 2 * ks, idesc, 1u);
          else tc_mma_i8(tmem, da + 2 * ks, db + 2 * ks, idesc, 1u);
        }
      }
      tc_commit(&bar[0]);
      mbar_wait(&bar[0], 0);
      if (blockIdx.x == 0) cyc[0] = clock64() - t0;
    }
  } else {
    const int w = warp - 1;
    const uint32_t base = tmem + (static_cast<uint32_t>((w & 3) * 32) << 16) + 256 + ((w >> 2) * 64) % 256;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < ld_iters; ++it) {
      uint32_t r[32];
      tmem_ld_32x32(base + (it & 1) * 32, r);
      tc_wait_ld();
      acc ^= r[0] ^ r[13] ^ r[31];
    }
    if (blockIdx.x == 0 && threadIdx.x == 32) cyc[1] = clock64() - t0;
  }
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int KIND>
static void run_ldtm_mma(int ld_warps, unsigned long long* d_cyc, uint32_t* d_sink) {
  const int smem = 16384 + 32768 + 1024 + 64;
  CK(cudaFuncSetAttribute(ldtm_mma_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int mma_iters = 20000, ld_iters = 20000;
  unsigned long long cyc[2] = {0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    ldtm_mma_kernel<KIND><<<148, 32 + ld_warps * 32, smem>>>(mma_iters, ld_iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
  printf("ldtm under kind::%s MMA, %2d reader warps: tcgen05.ld %6.1f B/clk per SM (alone: see above), MMA %6.1f cycles per instruction\n",
         KIND == 0 ? "f16" : "i8 ", ld_warps, static_cast<double>(ld_warps) * ld_iters * 4096.0 / cyc[1],
         static_cast<double>(cyc[0]) / (mma_iters * 4.0));
}

template <int DEPTH>
static void run_ldtm(int warps, int iters, unsigned long long* d_cyc, uint32_t* d_sink) {
  unsigned long long cyc = 0;
  for (int rep = 0; rep < 3; ++rep) {
    ldtm_kernel<DEPTH><<<148, warps * 32>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  const double bytes = static_cast<double>(warps) * iters * DEPTH * 32 * 32 * 4;
  printf("ldtm 32x32b.x32 warps=%2d depth=%d : %7.1f B/clk per SM (%llu cycles)\n", warps, DEPTH, bytes / cyc, cyc);
}

int main() {
  unsigned long long* d_cyc;
  uint32_t* d_sink;
  CK(cudaMalloc(&d_cyc, 64));
  CK(cudaMalloc(&d_sink, 4096));
  const int iters = 20000;
  for (int n : {256, 128, 64}) {
    run_mma<0, 1>(n, iters, d_cyc);
    run_mma<1, 1>(n, iters, d_cyc);
  }
  for (int n : {256, 128}) {
    run_mma<0, 2>(n, iters, d_cyc);
    run_mma<1, 2>(n, iters, d_cyc);
  }
  for (int warps : {4, 8, 16}) {
    run_ldtm<1>(warps, 20000, d_cyc, d_sink);
    run_ldtm<2>(warps, 20000, d_cyc, d_sink);
  }
  for (int warps : {8, 16}) {
    run_ldtm_mma<0>(warps, d_cyc, d_sink);
    run_ldtm_mma<1>(warps, d_cyc, d_sink);
  }
  return 0;
}
