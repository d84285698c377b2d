// Sampler step around the quantized UNet, fused into ONE elementwise kernel per denoising step:
//   classifier-free-guidance combine      pipeline_stable_diffusion.py:1037-1040
//   PLMS linear-multistep update          schedulers/scheduling_pndm.py:321-387, _get_prev_sample :407-449
//   Euler-ancestral update                schedulers/scheduling_euler_ancestral_discrete.py:323-414
//   next model input (scale + CFG concat) pipeline_stable_diffusion.py:1022-1024, scale_model_input
// Every one of these is linear in (sample, model outputs, noise); the host (dgq_b200/sampler.py) turns the
// scheduler state into coefficients, the kernel makes a single pass: HBM-bound, 16-byte vectorised.
#include "common.cuh"

namespace dgq {

struct SamplerDev {
  const float* unet_out;
  int64_t n;
  float guidance;
  int use_cfg;
  float* eps_store;
  const float* x;
  float cx, c_eps;
  const float* hist[4];
  float c_hist[4];
  const float* noise;
  float c_noise;
  float* out;
  float* model_in;
  float in_scale;
  int dup;
};

__device__ __forceinline__ float4 ld4(const float* p, int64_t i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }
__device__ __forceinline__ float4 fma4(float a, const float4 v, const float4 acc) {
  return make_float4(fmaf(a, v.x, acc.x), fmaf(a, v.y, acc.y), fmaf(a, v.z, acc.z), fmaf(a, v.w, acc.w));
}

__global__ void __launch_bounds__(256) sampler_step_kernel(const SamplerDev p) {
  const int64_t nvec = p.n >> 2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 e = ld4(p.unet_out, i);
    if (p.use_cfg) {
      const float4 c = ld4(p.unet_out + p.n, i);
      // noise_pred = uncond + g * (text - uncond): three separately rounded eager ops in the reference,
      // so no FMA contraction here (bit-exact eps)
      e = make_float4(__fadd_rn(e.x, __fmul_rn(p.guidance, __fsub_rn(c.x, e.x))),
                      __fadd_rn(e.y, __fmul_rn(p.guidance, __fsub_rn(c.y, e.y))),
                      __fadd_rn(e.z, __fmul_rn(p.guidance, __fsub_rn(c.z, e.z))),
                      __fadd_rn(e.w, __fmul_rn(p.guidance, __fsub_rn(c.w, e.w))));
    }
    if (p.eps_store != nullptr) reinterpret_cast<float4*>(p.eps_store)[i] = e;
    const float4 x = ld4(p.x, i);
    float4 acc = make_float4(p.cx * x.x, p.cx * x.y, p.cx * x.z, p.cx * x.w);
    acc = fma4(p.c_eps, e, acc);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (p.hist[k] != nullptr) acc = fma4(p.c_hist[k], ld4(p.hist[k], i), acc);
    }
    if (p.noise != nullptr) acc = fma4(p.c_noise, ld4(p.noise, i), acc);
    reinterpret_cast<float4*>(p.out)[i] = acc;
    if (p.model_in != nullptr) {
      const float4 m = make_float4(acc.x * p.in_scale, acc.y * p.in_scale, acc.z * p.in_scale, acc.w * p.in_scale);
      reinterpret_cast<float4*>(p.model_in)[i] = m;
      if (p.dup) reinterpret_cast<float4*>(p.model_in + p.n)[i] = m;
    }
  }
}

}  // namespace dgq

extern "C" int dgq_sampler_step(const dgq_sampler_step_t* a, void* stream) {
  using namespace dgq;
  DGQ_CHECK_ARG(a != nullptr && a->unet_out != nullptr && a->x != nullptr && a->out != nullptr);
  DGQ_CHECK_ARG(a->n > 0 && a->n % 4 == 0);
  DGQ_CHECK_ARG(a->c_noise == 0.0f || a->noise != nullptr);
  SamplerDev p;
  p.unet_out = a->unet_out; p.n = a->n; p.guidance = a->guidance; p.use_cfg = a->use_cfg;
  p.eps_store = a->eps_store; p.x = a->x; p.cx = a->cx; p.c_eps = a->c_eps;
  for (int k = 0; k < 4; ++k) { p.hist[k] = a->hist[k]; p.c_hist[k] = a->c_hist[k]; }
  p.noise = a->c_noise != 0.0f ? a->noise : nullptr; p.c_noise = a->c_noise;
  p.out = a->out; p.model_in = a->model_in; p.in_scale = a->in_scale; p.dup = a->dup;
  const int64_t nvec = a->n / 4;
  int64_t g = (nvec + 255) / 256;
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  sampler_step_kernel<<<static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DGQ_RETURN_LAST_ERROR();
}
