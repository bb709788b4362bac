// SAG decoder: Decoder_TRANSFORMER.forward of scripts/model/motionclip_module.py:137-183 -
// the step that turns a CLIP text feature z into the init_image of LivelySpeaker sampling
// (scripts/test_LivelySpeaker_ted.py:80-113).  3 post-norm nn.TransformerDecoderLayer (4 heads x 128, GELU FFN
// 512-1024-512) over 34 frame queries; the memory is ONE token, so the cross-attention softmax is over a single key
// and its output is out_proj(v_proj(z)) for every query.  Runs once per batch, before the diffusion loop.
//
// One CTA per clip, 512 threads, fp32 CUDA cores (this is 0.2 % of a T=1000 loop; the tensor-core treatment the
// denoiser got is not warranted).  Shared memory: X (tokens), Y (pre-norm sums), T (per-head Q/K/V or an FFN
// hidden half), all [34][512] fp32.  Every projection is "thread = output column, 34 accumulators in registers",
// reading the token rows as broadcast float4 and the TRANSPOSED weights ([k][n], prepared by the host mirror)
// coalesced; out-projection and FFN-2 accumulate across heads / hidden halves in registers.
#include <cstddef>

#include "ls_internal.cuh"

static_assert(sizeof(ls_sag_layer) == 144 && sizeof(ls_sag_weights) == 1232 && offsetof(ls_sag_weights, layer) == 80,
              "ls_sag_weights layout is part of the ABI (livelyspeaker_b200/sag.py mirrors it)");

namespace {

constexpr int D = 512, T = 34, HD = 128, NH = 4, FF = 1024;
constexpr int OFF_X = 0, OFF_Y = T * D, OFF_T = 2 * T * D, OFF_S = 3 * T * D, OFF_V = OFF_S + T * 36;
constexpr int SMEM_FLOATS = OFF_V + 2 * D;
constexpr float LN_EPS = 1e-5f;

// rows of src (+ nothing) -> LayerNorm -> dst; one warp per row
__device__ __forceinline__ void ln_rows(const float* src, float* dst, const float* __restrict__ g, const float* __restrict__ bta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < T; r += 16) {
    float v[16], s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] = src[r * D + lane + 32 * i];
      s += v[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.f / sqrtf(q * (1.f / D) + LN_EPS);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;
      dst[r * D + c] = (v[i] - mean) * rstd * g[c] + bta[c];
    }
  }
}

// acc[t] += sum_{k<K} in[t*ld + k] * wt[k*n + col]   (in: shared, rows of stride ld; wt: global, transposed weights)
template <int K>
__device__ __forceinline__ void gemm_col(const float* in, int ld, const float* __restrict__ wt, int n, int col, float (&acc)[T]) {
  const float* wc = wt + col;
#pragma unroll 1
  for (int k = 0; k < K; k += 4) {
    const float w0 = wc[(size_t)(k + 0) * n], w1 = wc[(size_t)(k + 1) * n], w2 = wc[(size_t)(k + 2) * n],
                w3 = wc[(size_t)(k + 3) * n];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float4 u = *reinterpret_cast<const float4*>(in + t * ld + k);
      acc[t] = fmaf(u.x, w0, acc[t]);
      acc[t] = fmaf(u.y, w1, acc[t]);
      acc[t] = fmaf(u.z, w2, acc[t]);
      acc[t] = fmaf(u.w, w3, acc[t]);
    }
  }
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(512, 1) sag_decode_kernel(ls_sag_weights w, const float* __restrict__ x,
                                                            const float* __restrict__ z, const uint8_t* __restrict__ mask,
                                                            float* __restrict__ out) {
  extern __shared__ float sm[];
  float* X = sm + OFF_X;
  float* Y = sm + OFF_Y;
  float* Tb = sm + OFF_T;
  float* S = sm + OFF_S;      // [34][36] attention probabilities of one head
  float* zv = sm + OFF_V;     // [512] z, then [512] v_proj(z)
  const int b = blockIdx.x, c = threadIdx.x;
  const int JD = w.njoints * w.nfeats;

  // ---- queries: mapping([motion of the first n_pre_poses frames, 1]) + positional encoding ----------------
  {
    const float* xb = x + (size_t)b * JD * T;
    const float bm = w.map_b[c];
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      float v = bm + w.pe[(size_t)t * w.pe_stride + c];
      if (t < w.n_pre_poses) {
        float a = w.map_wt[(size_t)JD * D + c];                    // the indicator bit
        for (int j = 0; j < JD; ++j) a = fmaf(xb[j * T + t], w.map_wt[(size_t)j * D + c], a);
        v += a;
      }
      X[t * D + c] = v;
    }
    zv[c] = z[(size_t)b * D + c];
  }
  __syncthreads();

  for (int l = 0; l < w.n_layers; ++l) {
    const ls_sag_layer& L = w.layer[l];
    // ---- self-attention ---------------------------------------------------------------------------------
    float oacc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) oacc[t] = 0.f;
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      if (c < 3 * HD) {             // q / k / v columns of this head: T[which][t][e]
        const int which = c / HD, e = c - which * HD, col = which * D + h * HD + e;
        float acc[T];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = 0.f;
        gemm_col<D>(X, D, L.sa_in_wt, 3 * D, col, acc);
        const float bq = L.sa_in_b[col], sc = which == 0 ? 0.08838834764831845f : 1.f;     // 1/sqrt(128)
#pragma unroll
        for (int t = 0; t < T; ++t) Tb[(which * T + t) * HD + e] = (acc[t] + bq) * sc;
      }
      __syncthreads();
      const float* Q = Tb;
      const float* K = Tb + T * HD;
      const float* V = Tb + 2 * T * HD;
      for (int i = c; i < T * T; i += 512) {      // scores
        const int t = i / T, s = i - t * T;
        float a = 0.f;
#pragma unroll 8
        for (int e = 0; e < HD; ++e) a = fmaf(Q[t * HD + e], K[s * HD + e], a);
        S[t * 36 + s] = a;
      }
      __syncthreads();
      if (c < T) {                                // softmax over the keys of row c
        float m = -INFINITY;
        for (int s = 0; s < T; ++s) m = fmaxf(m, S[c * 36 + s]);
        float sum = 0.f;
        for (int s = 0; s < T; ++s) {
          const float p = expf(S[c * 36 + s] - m);
          S[c * 36 + s] = p;
          sum += p;
        }
        const float inv = 1.f / sum;
        for (int s = 0; s < T; ++s) S[c * 36 + s] *= inv;
      }
      __syncthreads();
      {                                           // O[t][e] = sum_s P[t][s] V[s][e]  -> over the Q region
        const int e = c & (HD - 1), t0 = c >> 7;
        float o[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = 0.f;
        for (int s = 0; s < T; ++s) {
          const float v = V[s * HD + e];
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const int t = t0 + 4 * i;
            if (t < T) o[i] = fmaf(S[t * 36 + s], v, o[i]);
          }
        }
        __syncthreads();                          // every thread has finished reading Q (scores) before it is overwritten
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const int t = t0 + 4 * i;
          if (t < T) Tb[t * HD + e] = o[i];
        }
      }
      __syncthreads();
      gemm_col<HD>(Tb, HD, L.sa_out_wt + (size_t)h * HD * D, D, c, oacc);   // out_proj rows h*128 .. +127 of W^T
      __syncthreads();
    }
    {
      const float bo = L.sa_out_b[c];
#pragma unroll
      for (int t = 0; t < T; ++t) Y[t * D + c] = X[t * D + c] + oacc[t] + bo;
    }
    __syncthreads();
    ln_rows(Y, X, L.n1_w, L.n1_b);
    // ---- cross-attention over the single memory token -------------------------------------------------------
    {
      float a = L.ca_v_b[c];
      for (int k = 0; k < D; ++k) a = fmaf(zv[k], L.ca_v_wt[(size_t)k * D + c], a);
      zv[D + c] = a;
    }
    __syncthreads();
    {
      float a = L.ca_out_b[c];
      for (int k = 0; k < D; ++k) a = fmaf(zv[D + k], L.ca_out_wt[(size_t)k * D + c], a);
#pragma unroll
      for (int t = 0; t < T; ++t) Y[t * D + c] = X[t * D + c] + a;
    }
    __syncthreads();
    ln_rows(Y, X, L.n2_w, L.n2_b);
    __syncthreads();
    // ---- feed-forward: two hidden halves of 512 ---------------------------------------------------------------
#pragma unroll
    for (int t = 0; t < T; ++t) oacc[t] = 0.f;
#pragma unroll 1
    for (int half = 0; half < FF / D; ++half) {
      {
        float acc[T];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = 0.f;
        gemm_col<D>(X, D, L.l1_wt, FF, half * D + c, acc);
        const float b1 = L.l1_b[half * D + c];
#pragma unroll
        for (int t = 0; t < T; ++t) Tb[t * D + c] = gelu_exact(acc[t] + b1);
      }
      __syncthreads();
      gemm_col<D>(Tb, D, L.l2_wt + (size_t)half * D * D, D, c, oacc);
      __syncthreads();
    }
    {
      const float b2 = L.l2_b[c];
#pragma unroll
      for (int t = 0; t < T; ++t) Y[t * D + c] = X[t * D + c] + oacc[t] + b2;
    }
    __syncthreads();
    ln_rows(Y, X, L.n3_w, L.n3_b);
    __syncthreads();
  }
  // ---- final layer + padding mask; out [B, J*D, F] ---------------------------------------------------------------
  for (int i = c; i < JD * T; i += 512) {
    const int j = i / T, t = i - j * T;
    float a = w.fin_b[j];
    for (int k = 0; k < D; ++k) a = fmaf(X[t * D + k], w.fin_wt[(size_t)k * JD + j], a);
    if (mask != nullptr && !mask[(size_t)b * T + t]) a = 0.f;
    out[(size_t)b * JD * T + i] = a;
  }
}

}  // namespace

extern "C" int ls_sag_decode(const ls_sag_weights* w, int32_t B, const float* x, const float* z, const uint8_t* mask,
                             float* out, void* stream) {
  if (!w || !x || !z || !out || B < 1) return ls_fail(nullptr, LS_EINVAL, "ls_sag_decode: bad argument");
  if (w->n_frames != T || w->latent_dim != D || w->ff_size != FF || w->n_heads != NH || w->n_layers < 1 ||
      w->n_layers > LS_SAG_MAX_LAYERS || w->n_pre_poses < 0 || w->n_pre_poses > T)
    return ls_fail(nullptr, LS_EUNSUPPORTED, "ls_sag_decode: built for 34 frames, d=512, ff=1024, 4 heads, <= %d layers",
                   LS_SAG_MAX_LAYERS);
  const int smem = SMEM_FLOATS * (int)sizeof(float);
  // the attribute is per DEVICE: set on every launch (one process may drive several GPUs)
  if (cudaFuncSetAttribute(sag_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return ls_fail(nullptr, LS_ECUDA, "ls_sag_decode: cannot reserve %d bytes of shared memory", smem);
  sag_decode_kernel<<<B, 512, smem, (cudaStream_t)stream>>>(*w, x, z, mask, out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_sag_decode: %s", cudaGetErrorString(e));
  return LS_OK;
}
