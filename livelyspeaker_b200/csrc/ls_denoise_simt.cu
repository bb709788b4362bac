// fp32 CUDA-core denoiser: RAG.forward minus the hoisted terms, one CTA per
// (clip, pass).  This is the exact-order device implementation behind
// ls_model_forward / LS_IMPL_SIMT; the tcgen05 kernel in ls_fused.cu is the fast path.
//
//   input packing + input_mapping      scripts/model/RAG.py:110-114, 184-192
//   style token (reparameterize)       scripts/model/RAG.py:10-13, 117-122
//   TransMLP / MLPblock / LN_spatial   scripts/model/mlp_module.py:21-35, 67-74, 85-91
//   OutputProcess                      scripts/model/RAG.py:205-211
#include "ls_internal.cuh"

__device__ __forceinline__ float silu_f(float a) { return a / (1.f + expf(-a)); }

// LayerNorm over the 512 channels of every row of h -> u.  Warp per row, two-pass
// variance exactly like LN_spatial (mean, then mean of squared deviations).
template <int S>
__device__ __forceinline__ void ln_rows(const float* __restrict__ hs, float* __restrict__ us,
                                        const float* __restrict__ alpha, const float* __restrict__ beta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < S; r += 16) {
    float v[16];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] = hs[r * LS_D + lane + 32 * i];
      s += v[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / LS_D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float stdv = sqrtf(q * (1.f / LS_D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int c = lane + 32 * i;
      us[r * LS_D + c] = (v[i] - mean) / stdv * alpha[c] + beta[c];
    }
  }
}

template <int S>
__global__ void __launch_bounds__(512, 1)
denoise_simt_kernel(LsWeights w, int JD, int n_layers, int pass_mask, const float* __restrict__ x,
                    const int64_t* __restrict__ t, int t_uniform, const float* __restrict__ A,
                    const float* __restrict__ P, const float* __restrict__ z_mu, const float* __restrict__ z_lv,
                    const float* __restrict__ emo_tok, const float* __restrict__ eps_c,
                    const float* __restrict__ eps_u, float* __restrict__ out_c, float* __restrict__ out_u,
                    float* __restrict__ dbg_h, int dbg_layer, float* __restrict__ ckpt,
                    const uint8_t* __restrict__ cond_drop) {
  constexpr int NPRE = S - LS_F;
  extern __shared__ float sm[];
  float* hs = sm;                 // [S][512] residual stream
  float* us = hs + S * LS_D;      // [S][512] LN output / staging
  float* wt = us + S * LS_D;      // [S][S] token-mix weight, then [S] bias
  const int b = blockIdx.x, c = threadIdx.x;
  // cond_drop (training-mode mask_cond, RAG.py:84-93): clips whose audio embedding is replaced by zeros
  const bool uncond = ((pass_mask == 3) ? (blockIdx.y == 1) : (pass_mask == 2)) || (cond_drop != nullptr && cond_drop[b] != 0);
  const float* eps = uncond ? eps_u : eps_c;
  float* out = uncond ? out_u : out_c;

  // ---- tokens: style (, emotion), then 34 frames through the input projection ----
  const float* xb = x + (size_t)b * JD * LS_F;
  for (int i = c; i < JD * LS_F; i += 512) us[i] = xb[i];
  hs[c] = z_mu[(size_t)b * LS_D + c] + eps[(size_t)b * LS_D + c] * expf(0.5f * z_lv[(size_t)b * LS_D + c]);
  if (NPRE == 2) hs[LS_D + c] = emo_tok[(size_t)b * LS_D + c];
  __syncthreads();
  {
    float acc[LS_F];
#pragma unroll
    for (int f = 0; f < LS_F; ++f) acc[f] = 0.f;
    for (int j = 0; j < JD; ++j) {
      const float wv = w.w_x_t[(size_t)j * LS_D + c];
#pragma unroll
      for (int f = 0; f < LS_F; ++f) acc[f] = fmaf(us[j * LS_F + f], wv, acc[f]);
    }
    const float* Pb = P + (size_t)b * LS_F * LS_D;
    const float* Ab = A + (size_t)b * LS_F * LS_D;
#pragma unroll
    for (int f = 0; f < LS_F; ++f) {
      float v = acc[f] + Pb[f * LS_D + c];
      if (!uncond) v += Ab[f * LS_D + c];
      hs[(NPRE + f) * LS_D + c] = v;
    }
  }
  const long long tt = t_uniform >= 0 ? (long long)t_uniform : t[b];
  const float emb = w.emb_table[(size_t)tt * LS_D + c];
  __syncthreads();
  auto dump_hidden = [&](int l) {       // ls_debug_hidden: [b][pass][token][512]
    if (dbg_h != nullptr && dbg_layer == l)
      for (int r = 0; r < S; ++r) dbg_h[(((size_t)b * 2 + (uncond ? 1 : 0)) * S + r) * LS_D + c] = hs[r * LS_D + c];
  };
  dump_hidden(-1);

  for (int l = 0; l < n_layers; ++l) {
    const LsLayerW L = w.layer[l];
    if (ckpt != nullptr) {          // ls_cfg_forward_grad: the block's input is what the backward kernel restarts from
      float* ck = ckpt + (((size_t)b * 2 + (uncond ? 1 : 0)) * n_layers + l) * S * LS_D;
#pragma unroll
      for (int r = 0; r < S; ++r) ck[r * LS_D + c] = hs[r * LS_D + c];
    }
    // x = x + emb   (mlp_module.py:68-69; re-added in every block)
#pragma unroll
    for (int r = 0; r < S; ++r) hs[r * LS_D + c] += emb;
    for (int i = c; i < S * S; i += 512) wt[i] = L.w_tok[i];
    if (c < S) wt[S * S + c] = L.b_tok[c];
    __syncthreads();
    ln_rows<S>(hs, us, L.ln1_a, L.ln1_b);
    __syncthreads();
    // token mix: Conv1d(S,S,1) mixes rows, per channel (mlp_module.py:53)
    {
      float col[S];
#pragma unroll
      for (int j = 0; j < S; ++j) col[j] = us[j * LS_D + c];
#pragma unroll 1
      for (int i = 0; i < S; ++i) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < S; ++j) a = fmaf(wt[i * S + j], col[j], a);
        a += wt[S * S + i];
        hs[i * LS_D + c] += silu_f(a);
      }
    }
    __syncthreads();
    ln_rows<S>(hs, us, L.ln2_a, L.ln2_b);
    __syncthreads();
    // channel mix: Linear(512,512) (mlp_module.py:58); thread = output channel
    {
      float acc[S];
#pragma unroll
      for (int r = 0; r < S; ++r) acc[r] = 0.f;
      const float* wc = L.w_ch_t + c;
#pragma unroll 1
      for (int k = 0; k < LS_D; k += 4) {
        const float w0 = wc[(size_t)(k + 0) * LS_D], w1 = wc[(size_t)(k + 1) * LS_D],
                    w2 = wc[(size_t)(k + 2) * LS_D], w3 = wc[(size_t)(k + 3) * LS_D];
#pragma unroll
        for (int r = 0; r < S; ++r) {
          const float4 u4 = *reinterpret_cast<const float4*>(us + r * LS_D + k);
          acc[r] = fmaf(u4.x, w0, acc[r]);
          acc[r] = fmaf(u4.y, w1, acc[r]);
          acc[r] = fmaf(u4.z, w2, acc[r]);
          acc[r] = fmaf(u4.w, w3, acc[r]);
        }
      }
      const float bc = L.b_ch[c];
#pragma unroll
      for (int r = 0; r < S; ++r) hs[r * LS_D + c] += silu_f(acc[r] + bc);
    }
    dump_hidden(l);
    __syncthreads();
  }

  // ---- OutputProcess: drop the prefix tokens, Linear(512 -> JD), [B,J,D,F] layout ----
  const int warp = c >> 5, lane = c & 31;
  float* ob = out + (size_t)b * JD * LS_F;
  for (int o = warp; o < JD * LS_F; o += 16) {
    const int j = o / LS_F, f = o - j * LS_F;
    const float* hr = hs + (NPRE + f) * LS_D;
    const float* wr = w.w_out + (size_t)j * LS_D;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a = fmaf(hr[lane + 32 * i], wr[lane + 32 * i], a);
#pragma unroll
    for (int s = 16; s; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) ob[o] = a + w.b_out[j];
  }
}

template <int S>
static int launch_simt(ls_handle* h, int B, const float* x, const int64_t* t, int t_uniform, int pass_mask,
                       const float* eps_c, const float* eps_u, float* out_c, float* out_u, cudaStream_t s, float* ckpt,
                       const uint8_t* cond_drop) {
  const size_t smem = (size_t)(2 * S * LS_D + S * S + S) * sizeof(float);
  // the attribute is per DEVICE: set on every launch (one process may drive several GPUs)
  LS_CUDA(h, cudaFuncSetAttribute(denoise_simt_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(B, pass_mask == 3 ? 2 : 1);
  denoise_simt_kernel<S><<<grid, 512, smem, s>>>(h->w, h->JD, h->cfg.n_layers, pass_mask, x, t, t_uniform, h->A, h->P,
                                                 h->z_mu, h->z_lv, h->emo_tok, eps_c, eps_u, out_c, out_u, h->dbg_h,
                                                 h->dbg_layer, ckpt, cond_drop);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

int lsk_denoise_simt(ls_handle* h, int B, const float* x, const int64_t* t, int t_uniform, int pass_mask,
                     const float* eps_c, const float* eps_u, float* out_c, float* out_u, cudaStream_t s, float* ckpt,
                     const uint8_t* cond_drop) {
  if (h->S == 35) return launch_simt<35>(h, B, x, t, t_uniform, pass_mask, eps_c, eps_u, out_c, out_u, s, ckpt, cond_drop);
  if (h->S == 36) return launch_simt<36>(h, B, x, t, t_uniform, pass_mask, eps_c, eps_u, out_c, out_u, s, ckpt, cond_drop);
  return ls_fail(h, LS_EUNSUPPORTED, "token count %d not built", h->S);
}

// ------------------------------------------------------------------------------------------------------------------
// Backward of the denoiser with respect to x (vector-Jacobian product): what torch autograd does under
// p_sample_with_grad / ddim_sample_with_grad / condition_*_with_grad (gaussian_diffusion.py:444-505, 560-606, 800-855),
// where cond_fn differentiates a function of p_mean_var['pred_xstart'] with respect to x.  One CTA per (clip, pass),
// thread = channel.  The forward kernel leaves the INPUT of every MLPblock in `ckpt`; the backward walks the blocks in
// reverse, recomputes a block's forward from its checkpoint and backpropagates through it with three [S][512] tiles
// in shared memory:
//   Bg  gradient (dL/d block output on entry, dL/d block input on exit)
//   Bh  a = h + emb, then m = a + silu(token mix)
//   Bu  y (token-mix pre-activation), then LN1's normalised input
// LayerNorm 2 is folded into the recomputed channel mix (z = rho (m . (alpha W) - mu S) + T) so that no fourth tile is
// needed.  MLPblock: mlp_module.py:67-74; LN_spatial: :29-35; input / output processes: RAG.py:184-211.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dsilu_f(float a) {
  const float sg = 1.f / (1.f + expf(-a));
  return sg * (1.f + a * (1.f - sg));
}

struct LsRawCh { const float* w[LS_MAX_LAYERS]; };      // block2.1.weight [out][in], untransposed

// per row: mean and 1 / sqrt(var + eps) over the 512 channels (two-pass, like ln_rows)
template <int S>
__device__ __forceinline__ void row_stats(const float* __restrict__ hs, float* mu, float* rho) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < S; r += 16) {
    float v[16], s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] = hs[r * LS_D + lane + 32 * i];
      s += v[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / LS_D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) {
      mu[r] = mean;
      rho[r] = 1.f / sqrtf(q * (1.f / LS_D) + 1e-5f);
    }
  }
}
// per row: c1 = mean_c g, c2 = mean_c g * uhat;  uhat = NORM ? (hs - mu) * rho : hs
template <int S, bool NORM>
__device__ __forceinline__ void row_dots(const float* __restrict__ g, const float* __restrict__ hs, const float* mu,
                                         const float* rho, float* c1, float* c2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < S; r += 16) {
    float s1 = 0.f, s2 = 0.f;
    const float m = NORM ? mu[r] : 0.f, rh = NORM ? rho[r] : 1.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float gv = g[r * LS_D + lane + 32 * i], uh = (hs[r * LS_D + lane + 32 * i] - m) * rh;
      s1 += gv;
      s2 = fmaf(gv, uh, s2);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      c1[r] = s1 * (1.f / LS_D);
      c2[r] = s2 * (1.f / LS_D);
    }
  }
}

template <int S>
__global__ void __launch_bounds__(512, 1)
denoise_simt_bwd_kernel(LsWeights w, LsRawCh raw, int JD, int n_layers, int B, const int64_t* __restrict__ t,
                        const float* __restrict__ ckpt, const float* __restrict__ grad_out,
                        const float* __restrict__ scale, float* __restrict__ gx) {
  constexpr int NPRE = S - LS_F;
  extern __shared__ float sm[];
  float* Bg = sm;
  float* Bh = Bg + S * LS_D;
  float* Bu = Bh + S * LS_D;
  float* wt = Bu + S * LS_D;        // [S][S] token-mix weight, [S] bias
  float* mu1 = wt + S * S + S;      // 6 x [S] per-row scalars
  float* rho1 = mu1 + S;
  float* mu2 = rho1 + S;
  float* rho2 = mu2 + S;
  float* c1 = rho2 + S;
  float* c2 = c1 + S;
  const int b = blockIdx.x, pass = blockIdx.y, c = threadIdx.x;
  // out = out_u + s (out_c - out_u)  (cfg_sampler.py:31): d out / d out_c = s, d out / d out_u = 1 - s
  const float fac = pass == 0 ? scale[b] : 1.f - scale[b];

  // ---- OutputProcess backward: g[k+f][c] = sum_j go[j][f] * W_out[j][c]; prefix tokens receive nothing ----
  for (int i = c; i < JD * LS_F; i += 512) Bu[i] = fac * grad_out[(size_t)b * JD * LS_F + i];
  __syncthreads();
  {
    float acc[LS_F];
#pragma unroll
    for (int f = 0; f < LS_F; ++f) acc[f] = 0.f;
    for (int j = 0; j < JD; ++j) {
      const float wv = w.w_out[(size_t)j * LS_D + c];
#pragma unroll
      for (int f = 0; f < LS_F; ++f) acc[f] = fmaf(Bu[j * LS_F + f], wv, acc[f]);
    }
#pragma unroll
    for (int r = 0; r < NPRE; ++r) Bg[r * LS_D + c] = 0.f;
#pragma unroll
    for (int f = 0; f < LS_F; ++f) Bg[(NPRE + f) * LS_D + c] = acc[f];
  }
  const float emb = w.emb_table[(size_t)t[b] * LS_D + c];
  __syncthreads();

  for (int l = n_layers - 1; l >= 0; --l) {
    const LsLayerW L = w.layer[l];
    const float* ck = ckpt + (((size_t)b * 2 + pass) * n_layers + l) * S * LS_D;
    const float a1 = L.ln1_a[c], b1 = L.ln1_b[c], a2 = L.ln2_a[c];
    // ---- recompute the block's forward: a, LN1, token mix ----
#pragma unroll
    for (int r = 0; r < S; ++r) Bh[r * LS_D + c] = ck[r * LS_D + c] + emb;
    for (int i = c; i < S * S; i += 512) wt[i] = L.w_tok[i];
    if (c < S) wt[S * S + c] = L.b_tok[c];
    __syncthreads();
    row_stats<S>(Bh, mu1, rho1);
    __syncthreads();
    {
      float col[S];
#pragma unroll
      for (int j = 0; j < S; ++j) col[j] = (Bh[j * LS_D + c] - mu1[j]) * rho1[j] * a1 + b1;
#pragma unroll 1
      for (int i = 0; i < S; ++i) {
        float y = 0.f;
#pragma unroll
        for (int j = 0; j < S; ++j) y = fmaf(wt[i * S + j], col[j], y);
        y += wt[S * S + i];
        Bu[i * LS_D + c] = y;
        Bh[i * LS_D + c] += silu_f(y);           // m
      }
    }
    __syncthreads();
    row_stats<S>(Bh, mu2, rho2);
    __syncthreads();
    // ---- channel mix recomputed with LayerNorm 2 folded in; gz = g * silu'(z) ----
    float gcol[S];
    {
      float acc[S];
#pragma unroll
      for (int r = 0; r < S; ++r) acc[r] = 0.f;
      float Ssum = 0.f, Tsum = 0.f;
      const float* wc = L.w_ch_t + c;
#pragma unroll 1
      for (int k = 0; k < LS_D; k += 4) {
        const float4 al = *reinterpret_cast<const float4*>(L.ln2_a + k), be = *reinterpret_cast<const float4*>(L.ln2_b + k);
        const float r0 = wc[(size_t)(k + 0) * LS_D], r1 = wc[(size_t)(k + 1) * LS_D], r2 = wc[(size_t)(k + 2) * LS_D],
                    r3 = wc[(size_t)(k + 3) * LS_D];
        const float w0 = r0 * al.x, w1 = r1 * al.y, w2 = r2 * al.z, w3 = r3 * al.w;
        Ssum += (w0 + w1) + (w2 + w3);
        Tsum = fmaf(r0, be.x, fmaf(r1, be.y, fmaf(r2, be.z, fmaf(r3, be.w, Tsum))));
#pragma unroll
        for (int r = 0; r < S; ++r) {
          const float4 u4 = *reinterpret_cast<const float4*>(Bh + r * LS_D + k);
          acc[r] = fmaf(u4.x, w0, acc[r]);
          acc[r] = fmaf(u4.y, w1, acc[r]);
          acc[r] = fmaf(u4.z, w2, acc[r]);
          acc[r] = fmaf(u4.w, w3, acc[r]);
        }
      }
      const float bc = L.b_ch[c];
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const float z = rho2[r] * (acc[r] - mu2[r] * Ssum) + Tsum + bc;
        gcol[r] = Bg[r * LS_D + c];
        Bg[r * LS_D + c] = gcol[r] * dsilu_f(z);   // own column only: nobody reads Bg rows before the barrier
      }
    }
    __syncthreads();
    // ---- d/d u2: gu2[r][c] = sum_c' gz[r][c'] W[c'][c]; times alpha2 = gradient of LN2's normalised input ----
    {
      float acc[S];
#pragma unroll
      for (int r = 0; r < S; ++r) acc[r] = 0.f;
      const float* wc = raw.w[l] + c;
#pragma unroll 1
      for (int k = 0; k < LS_D; k += 4) {
        const float w0 = wc[(size_t)(k + 0) * LS_D], w1 = wc[(size_t)(k + 1) * LS_D], w2 = wc[(size_t)(k + 2) * LS_D],
                    w3 = wc[(size_t)(k + 3) * LS_D];
#pragma unroll
        for (int r = 0; r < S; ++r) {
          const float4 u4 = *reinterpret_cast<const float4*>(Bg + r * LS_D + k);
          acc[r] = fmaf(u4.x, w0, acc[r]);
          acc[r] = fmaf(u4.y, w1, acc[r]);
          acc[r] = fmaf(u4.z, w2, acc[r]);
          acc[r] = fmaf(u4.w, w3, acc[r]);
        }
      }
      __syncthreads();                           // every thread has finished reading the rows of gz
#pragma unroll
      for (int r = 0; r < S; ++r) Bg[r * LS_D + c] = acc[r] * a2;
    }
    __syncthreads();
    row_dots<S, true>(Bg, Bh, mu2, rho2, c1, c2);
    __syncthreads();
    // ---- LN2 backward -> gm; token mix backward -> gradient of LN1's normalised input ----
    {
      float gy[S];
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const float uh = (Bh[r * LS_D + c] - mu2[r]) * rho2[r];
        gcol[r] += rho2[r] * (Bg[r * LS_D + c] - c1[r] - uh * c2[r]);      // gm
        gy[r] = gcol[r] * dsilu_f(Bu[r * LS_D + c]);
      }
      __syncthreads();                           // row_dots' readers are done with Bg / Bh rows; Bu is column-private so far
#pragma unroll 1
      for (int j = 0; j < S; ++j) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < S; ++i) a = fmaf(wt[i * S + j], gy[i], a);
        const float y = Bu[j * LS_D + c];
        const float av = Bh[j * LS_D + c] - silu_f(y);                     // a = m - silu(y)
        Bg[j * LS_D + c] = a * a1;
        Bu[j * LS_D + c] = (av - mu1[j]) * rho1[j];                        // LN1's normalised input
      }
    }
    __syncthreads();
    row_dots<S, false>(Bg, Bu, mu1, rho1, c1, c2);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < S; ++r)
      gcol[r] += rho1[r] * (Bg[r * LS_D + c] - c1[r] - Bu[r * LS_D + c] * c2[r]);                 // ga = dL/d block input
    __syncthreads();
#pragma unroll
    for (int r = 0; r < S; ++r) Bg[r * LS_D + c] = gcol[r];
    __syncthreads();
  }

  // ---- InputProcess backward: gx[j][f] = sum_c g[k+f][c] * W_x[c][j]  (RAG.py:184-192; the hoisted terms do not depend on x) ----
  const int warp = c >> 5, lane = c & 31;
  float* gxb = gx + ((size_t)pass * B + b) * JD * LS_F;
  for (int o = warp; o < JD * LS_F; o += 16) {
    const int j = o / LS_F, f = o - j * LS_F;
    const float* gr = Bg + (NPRE + f) * LS_D;
    const float* wr = w.w_x_t + (size_t)j * LS_D;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a = fmaf(gr[lane + 32 * i], wr[lane + 32 * i], a);
#pragma unroll
    for (int s2 = 16; s2; s2 >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s2);
    if (lane == 0) gxb[o] = a;
  }
}

template <int S>
static int launch_simt_bwd(ls_handle* h, int B, const int64_t* t, const float* ckpt, const float* grad_out,
                           const float* scale, float* gx, cudaStream_t s) {
  const size_t smem = (size_t)(3 * S * LS_D + S * S + S + 6 * S) * sizeof(float);
  LS_CUDA(h, cudaFuncSetAttribute(denoise_simt_bwd_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LsRawCh raw{};
  for (int l = 0; l < h->cfg.n_layers; ++l) {
    const std::string key = "backbone.mlps." + std::to_string(l) + ".block2.1.weight";
    for (auto& r : h->raw)
      if (r.key == key) raw.w[l] = r.dev;
    if (raw.w[l] == nullptr) return ls_fail(h, LS_ESTATE, "missing %s", key.c_str());
  }
  denoise_simt_bwd_kernel<S><<<dim3(B, 2), 512, smem, s>>>(h->w, raw, h->JD, h->cfg.n_layers, B, t, ckpt, grad_out, scale, gx);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

// gx: [2][B][JD][34] per-pass gradients (the caller adds the two)
int lsk_denoise_simt_bwd(ls_handle* h, int B, const int64_t* t, const float* ckpt, const float* grad_out,
                         const float* scale, float* gx, cudaStream_t s) {
  if (h->S == 35) return launch_simt_bwd<35>(h, B, t, ckpt, grad_out, scale, gx, s);
  if (h->S == 36) return launch_simt_bwd<36>(h, B, t, ckpt, grad_out, scale, gx, s);
  return ls_fail(h, LS_EUNSUPPORTED, "token count %d not built", h->S);
}
