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
                    float* __restrict__ dbg_h, int dbg_layer) {
  constexpr int NPRE = S - LS_F;
  extern __shared__ float sm[];
  float* hs = sm;                 // [S][512] residual stream
  float* us = hs + S * LS_D;      // [S][512] LN output / staging
  float* wt = us + S * LS_D;      // [S][S] token-mix weight, then [S] bias
  const int b = blockIdx.x, c = threadIdx.x;
  const bool uncond = (pass_mask == 3) ? (blockIdx.y == 1) : (pass_mask == 2);
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
                       const float* eps_c, const float* eps_u, float* out_c, float* out_u, cudaStream_t s) {
  const size_t smem = (size_t)(2 * S * LS_D + S * S + S) * sizeof(float);
  // the attribute is per DEVICE: set on every launch (one process may drive several GPUs)
  LS_CUDA(h, cudaFuncSetAttribute(denoise_simt_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(B, pass_mask == 3 ? 2 : 1);
  denoise_simt_kernel<S><<<grid, 512, smem, s>>>(h->w, h->JD, h->cfg.n_layers, pass_mask, x, t, t_uniform, h->A, h->P,
                                                 h->z_mu, h->z_lv, h->emo_tok, eps_c, eps_u, out_c, out_u, h->dbg_h,
                                                 h->dbg_layer);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

int lsk_denoise_simt(ls_handle* h, int B, const float* x, const int64_t* t, int t_uniform, int pass_mask,
                     const float* eps_c, const float* eps_u, float* out_c, float* out_u, cudaStream_t s) {
  if (h->S == 35) return launch_simt<35>(h, B, x, t, t_uniform, pass_mask, eps_c, eps_u, out_c, out_u, s);
  if (h->S == 36) return launch_simt<36>(h, B, x, t, t_uniform, pass_mask, eps_c, eps_u, out_c, out_u, s);
  return ls_fail(h, LS_EUNSUPPORTED, "token count %d not built", h->S);
}
