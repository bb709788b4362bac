// The per-element sampler update shared by the SIMT tail kernel and the fused
// tcgen05 kernel's epilogue.  Operation order follows the reference so fp32 results
// stay within an ulp or two of torch's:
//   gaussian_diffusion.py:268-271 (posterior mean), :557 (ancestral sample),
//   :418-422 (eps from x0), :790-797 (DDIM mean / sample), :365-371 (clip).
#pragma once
#include "../../include/livelyspeaker_b200.h"

// Returns the (possibly clipped) x0; writes x_{t-1} to *x_prev (untouched for mode 2).
// `noise` must already be 0 when p.add_noise == 0.
__device__ __forceinline__ float ls_sampler_update(const ls_step_params& p, float x0, float x_t, float noise,
                                                   float* x_prev) {
  if (p.clip_denoised) x0 = fminf(fmaxf(x0, -1.f), 1.f);
  if (p.mode == 0) {
    const float mean = __fadd_rn(__fmul_rn(p.c[0], x0), __fmul_rn(p.c[1], x_t));
    const float sd = p.add_noise ? expf(0.5f * p.c[2]) : 0.f;
    *x_prev = __fadd_rn(mean, __fmul_rn(sd, noise));
  } else if (p.mode == 1) {
    const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(p.c[0], x_t), x0), p.c[1]);
    const float mean = __fadd_rn(__fmul_rn(x0, p.c[2]), __fmul_rn(p.c[3], eps));
    const float sg = p.add_noise ? p.c[4] : 0.f;
    *x_prev = __fadd_rn(mean, __fmul_rn(sg, noise));
  }
  return x0;
}
