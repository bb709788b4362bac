// Elementwise tail of a denoising step for the non-fused (SIMT) path:
//   classifier-free combine     scripts/model/cfg_sampler.py:31
//   clip_denoised               scripts/diffusion/gaussian_diffusion.py:365-371
//   ancestral update            scripts/diffusion/gaussian_diffusion.py:260-282, 547-557
//   DDIM update                 scripts/diffusion/gaussian_diffusion.py:418-422, 777-797
//   q_sample                    scripts/diffusion/gaussian_diffusion.py:240-258
#include "ls_internal.cuh"
#include "ls_update.cuh"

__global__ void cfg_combine_kernel(int n_per, long long n, const float* __restrict__ oc, const float* __restrict__ ou,
                                   const float* __restrict__ scale, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = scale[i / n_per];
  out[i] = ou[i] + s * (oc[i] - ou[i]);
}

int lsk_cfg_combine(ls_handle* h, int B, const float* out_c, const float* out_u, const float* scale, float* out,
                    cudaStream_t s) {
  const int n_per = h->JD * LS_F;
  const long long n = (long long)B * n_per;
  cfg_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n_per, n, out_c, out_u, scale, out);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

__global__ void cfg_update_kernel(int JD, long long n, ls_step_params p, const float* __restrict__ oc,
                                  const float* __restrict__ ou, const float* __restrict__ scale,
                                  const float* __restrict__ x_t, const float* __restrict__ noise, long long sb,
                                  long long sj, long long sf, float* __restrict__ x_prev, float* __restrict__ pred_x0) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int n_per = JD * LS_F;
  const long long b = i / n_per;
  const int r = (int)(i - b * n_per), jd = r / LS_F, f = r - jd * LS_F;
  float x0 = ou[i] + scale[b] * (oc[i] - ou[i]);
  const float nz = (p.mode != 2 && p.add_noise) ? noise[b * sb + jd * sj + f * sf] : 0.f;
  float xp;
  x0 = ls_sampler_update(p, x0, x_t[i], nz, &xp);
  if (pred_x0) pred_x0[i] = x0;
  if (p.mode != 2) x_prev[i] = xp;
}

int lsk_cfg_update(ls_handle* h, int B, const ls_step_params* p, const float* out_c, const float* out_u,
                   const float* scale, const float* x_t, const float* noise, int64_t sb, int64_t sj, int64_t sf,
                   float* x_prev, float* pred_x0, cudaStream_t s) {
  const long long n = (long long)B * h->JD * LS_F;
  cfg_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->JD, n, *p, out_c, out_u, scale, x_t, noise, sb, sj, sf,
                                                              x_prev, pred_x0);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

__global__ void axpby_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b, float ca, float cb,
                             float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ca * a[i] + cb * b[i];
}

int lsk_axpby(ls_handle* h, int64_t n, const float* a, const float* b, float ca, float cb, float* out, cudaStream_t s) {
  axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, a, b, ca, cb, out);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

// scale[b] that makes the guided combination out_u + scale (out_c - out_u) return ONE pass: 1 = cond, 0 = uncond;
// per clip 1 - (drop[b] != 0) when a training-mode condition-dropout mask is given (RAG.py:84-93)
__global__ void pass_scale_kernel(int B, int cond, const uint8_t* __restrict__ drop, float* __restrict__ scale) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) scale[b] = drop != nullptr ? (drop[b] != 0 ? 0.f : 1.f) : (cond ? 1.f : 0.f);
}
int lsk_pass_scale(ls_handle* h, int B, int cond, const uint8_t* drop, float* scale, cudaStream_t s) {
  pass_scale_kernel<<<(B + 255) / 256, 256, 0, s>>>(B, cond, drop, scale);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}
