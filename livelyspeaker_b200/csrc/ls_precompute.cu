// Step-invariant part of RAG.forward, computed once per batch instead of once per
// model call (the reference recomputes all of it 2x per denoising step):
//   WavEncoder                      scripts/model/audio_enc.py:6-25
//   audio / prefix halves of input_mapping, speaker mu/logvar, emotion token
//                                   scripts/model/RAG.py:106-120, 184-192
//   time_embed(pe[t]) table         scripts/model/mlp_module.py:123-136
// All fp32 on CUDA cores: this is < 2 % of a T=1000 loop.
#include <cstdlib>

#include "ls_internal.cuh"

// ------------------------------------------------------------------------------------
// Conv1d, kernel 15, arbitrary stride/padding.  Block = 32 output channels x 64 output
// positions of one clip; thread = 4 channels x 4 positions (positions interleaved by
// 16 so the strided smem reads are conflict free for stride 5 and 6).
// ------------------------------------------------------------------------------------
constexpr int CV_TC = 32, CV_TL = 64, CV_CC = 8, CV_K = 15;

template <int STRIDE>
__global__ void __launch_bounds__(128) conv1d_k15_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         int Ci, int Li, int Co, int Lo, int pad) {
  constexpr int XW = (CV_TL - 1) * STRIDE + CV_K;
  __shared__ float xs[CV_CC][XW];
  __shared__ float ws[CV_TC * CV_CC * CV_K];
  const int b = blockIdx.z, co0 = blockIdx.y * CV_TC, lo0 = blockIdx.x * CV_TL;
  const int tid = threadIdx.x, tl = tid & 15, tc = tid >> 4;
  const float* inb = in + (size_t)b * Ci * Li;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  const int x0 = lo0 * STRIDE - pad;
  for (int c0 = 0; c0 < Ci; c0 += CV_CC) {
    const int cc = min(CV_CC, Ci - c0);
    for (int idx = tid; idx < cc * XW; idx += 128) {
      int ci = idx / XW, p = idx - ci * XW, g = x0 + p;
      xs[ci][p] = (g >= 0 && g < Li) ? inb[(size_t)(c0 + ci) * Li + g] : 0.f;
    }
    for (int idx = tid; idx < CV_TC * cc * CV_K; idx += 128) {
      int co = idx / (cc * CV_K), r = idx - co * (cc * CV_K);
      ws[co * (CV_CC * CV_K) + r] = (co0 + co < Co) ? w[((size_t)(co0 + co) * Ci + c0) * CV_K + r] : 0.f;
    }
    __syncthreads();
    for (int ci = 0; ci < cc; ++ci) {
#pragma unroll
      for (int k = 0; k < CV_K; ++k) {
        float xv[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = xs[ci][(tl + 16 * i) * STRIDE + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = ws[(tc * 4 + j) * (CV_CC * CV_K) + ci * CV_K + k];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(wv[j], xv[i], acc[j][i]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int co = co0 + tc * 4 + j;
    if (co >= Co) continue;
    float bv = bias[co];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int lo = lo0 + tl + 16 * i;
      if (lo < Lo) out[((size_t)b * Co + co) * Lo + lo] = acc[j][i] + bv;
    }
  }
}

// InstanceNorm1d (no affine, biased variance, eps 1e-5) + LeakyReLU(0.3), in place.
// One block per (clip, channel).  audio_enc.py:10-17.
__global__ void __launch_bounds__(256) instnorm_lrelu_kernel(float* __restrict__ x, int L) {
  __shared__ float red[8];
  __shared__ float stat;
  float* row = x + (size_t)blockIdx.x * L;
  const int tid = threadIdx.x;
  float s = 0.f;
  for (int i = tid; i < L; i += 256) s += row[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    stat = t / (float)L;
  }
  __syncthreads();
  const float mean = stat;
  float v = 0.f;
  for (int i = tid; i < L; i += 256) {
    float d = row[i] - mean;
    v = fmaf(d, d, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    stat = rsqrtf(t / (float)L + 1e-5f);
  }
  __syncthreads();
  const float inv = stat;
  for (int i = tid; i < L; i += 256) {
    float y = (row[i] - mean) * inv;
    row[i] = y > 0.f ? y : 0.3f * y;
  }
}

static int conv_out_len(int Li, int pad, int stride) { return (Li + 2 * pad - CV_K) / stride + 1; }

// audio [B, L] -> out_cm [B, 256, 34] (channel-major, i.e. the conv's own layout).
int lsk_wav_encoder(ls_handle* h, int B, const float* audio, float* out_cm, cudaStream_t s) {
  const int L0 = h->cfg.audio_len;
  const int L1 = conv_out_len(L0, 1600, 5), L2 = conv_out_len(L1, 0, 6), L3 = conv_out_len(L2, 0, 6),
            L4 = conv_out_len(L3, 0, 6);
  if (L4 != LS_F) return ls_fail(h, LS_EINVAL, "audio_len %d maps to %d frames, need %d", L0, L4, LS_F);
  auto W = [&](const char* k) -> const float* {
    for (auto& r : h->raw)
      if (r.key == k) return r.dev;
    return nullptr;
  };
  const float *w0 = W("audio_encoder.feat_extractor.0.weight"), *b0 = W("audio_encoder.feat_extractor.0.bias");
  const float *w1 = W("audio_encoder.feat_extractor.3.weight"), *b1 = W("audio_encoder.feat_extractor.3.bias");
  const float *w2 = W("audio_encoder.feat_extractor.6.weight"), *b2 = W("audio_encoder.feat_extractor.6.bias");
  const float *w3 = W("audio_encoder.feat_extractor.9.weight"), *b3 = W("audio_encoder.feat_extractor.9.bias");
  // layers 2-4 on the tensor cores (bf16x3) unless the exact-order fp32 implementation is selected
  const bool tc = lsw_available(h) && ls_get_impl(h) != LS_IMPL_SIMT;
  // LS_WAV_V1=1 (diagnostic): the round-1 pipeline (separate InstanceNorm passes) on the tensor-core path
  static const bool v1 = getenv("LS_WAV_V1") != nullptr;
  for (int c0 = 0; c0 < B; c0 += h->wav_chunk) {
    const int nb = min(h->wav_chunk, B - c0);
    const float* a = audio + (size_t)c0 * L0;
    if (tc && !v1) {
      const int rc = lsw_encoder_fused(h, a, w0, b3, out_cm + (size_t)c0 * LS_AF * LS_F, nb, L0, L1, L2, L3, L4, s);
      if (rc) return rc;
      continue;
    }
    conv1d_k15_kernel<5><<<dim3((L1 + CV_TL - 1) / CV_TL, 1, nb), 128, 0, s>>>(a, w0, b0, h->wav_a, 1, L0, 32, L1, 1600);
    LS_LAUNCH_CHECK(h);
    instnorm_lrelu_kernel<<<nb * 32, 256, 0, s>>>(h->wav_a, L1);
    LS_LAUNCH_CHECK(h);
    float* out4 = out_cm + (size_t)c0 * LS_AF * LS_F;
    int rc;
    if (tc) {
      if ((rc = lsw_conv(h, 0, h->wav_a, b1, h->wav_b, nb, L1, L2, s))) return rc;
    } else {
      conv1d_k15_kernel<6><<<dim3((L2 + CV_TL - 1) / CV_TL, 2, nb), 128, 0, s>>>(h->wav_a, w1, b1, h->wav_b, 32, L1, 64, L2, 0);
      LS_LAUNCH_CHECK(h);
    }
    instnorm_lrelu_kernel<<<nb * 64, 256, 0, s>>>(h->wav_b, L2);
    LS_LAUNCH_CHECK(h);
    if (tc) {
      if ((rc = lsw_conv(h, 1, h->wav_b, b2, h->wav_a, nb, L2, L3, s))) return rc;
    } else {
      conv1d_k15_kernel<6><<<dim3((L3 + CV_TL - 1) / CV_TL, 4, nb), 128, 0, s>>>(h->wav_b, w2, b2, h->wav_a, 64, L2, 128, L3, 0);
      LS_LAUNCH_CHECK(h);
    }
    instnorm_lrelu_kernel<<<nb * 128, 256, 0, s>>>(h->wav_a, L3);
    LS_LAUNCH_CHECK(h);
    if (tc) {
      if ((rc = lsw_conv(h, 2, h->wav_a, b3, out4, nb, L3, L4, s))) return rc;
    } else {
      conv1d_k15_kernel<6><<<dim3((L4 + CV_TL - 1) / CV_TL, 8, nb), 128, 0, s>>>(h->wav_a, w3, b3, out4, 128, L3, 256, L4, 0);
      LS_LAUNCH_CHECK(h);
    }
  }
  return LS_OK;
}

// [B,256,34] -> [B,34,256] (WavEncoder.forward's transpose, audio_enc.py:25)
__global__ void cm_to_fm_kernel(const float* __restrict__ in, float* __restrict__ out) {
  const float* a = in + (size_t)blockIdx.x * LS_AF * LS_F;
  float* o = out + (size_t)blockIdx.x * LS_AF * LS_F;
  for (int i = threadIdx.x; i < LS_AF * LS_F; i += blockDim.x) {
    int f = i / LS_AF, c = i - f * LS_AF;
    o[i] = a[c * LS_F + f];
  }
}
int lsk_cm_to_fm(ls_handle* h, const float* in_cm, float* out_fm, int B, cudaStream_t s) {
  cm_to_fm_kernel<<<B, 256, 0, s>>>(in_cm, out_fm);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

// ------------------------------------------------------------------------------------
// Hoisted conditioning.  One block per clip, thread = latent channel c.
//   A[b,f,c] = sum_k af[b,f,k] * W_in[c, 2JD+1+k]                      (cond pass only)
//   P[b,f,c] = b_in[c] + [f<4] * ( W_in[c,2JD] + sum_j ox[b,j,f] * W_in[c,JD+j] )
//   mu / logvar = Linear(speaker_embedding[vid])            (RAG.py:117-119)
//   emo_tok = emotion_embedding[emo[b,0]]                   (scripts_beat/model/RAG.py:125)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) cond_proj_kernel(LsWeights w, int JD, int n_spk, int n_emo,
                                                        const float* __restrict__ af_cm, float* __restrict__ origin_x,
                                                        const int64_t* __restrict__ vid, const int64_t* __restrict__ emo,
                                                        int64_t emo_stride, int mutate_origin, int b0, int with_A,
                                                        float* __restrict__ A, float* __restrict__ P,
                                                        float* __restrict__ z_mu, float* __restrict__ z_lv,
                                                        float* __restrict__ emo_tok) {
  extern __shared__ float sm[];
  float* af_s = sm;                       // [256][34]
  float* ox_s = af_s + LS_AF * LS_F;      // [JD][4]
  float* z_s = ox_s + JD * LS_NPRE;       // [256]
  const int b = b0 + blockIdx.x, c = threadIdx.x;
  const float* afb = af_cm + (size_t)blockIdx.x * LS_AF * LS_F;
  if (with_A)
    for (int i = c; i < LS_AF * LS_F; i += 512) af_s[i] = afb[i];
  float* oxb = origin_x + (size_t)b * JD * LS_F;
  for (int i = c; i < JD * LS_NPRE; i += 512) ox_s[i] = oxb[(i >> 2) * LS_F + (i & 3)];
  long long v = vid[b];
  v = v < 0 ? 0 : (v >= n_spk ? n_spk - 1 : v);
  if (c < LS_SPK) z_s[c] = w.spk_emb[(size_t)v * LS_SPK + c];
  __syncthreads();
  if (mutate_origin)
    for (int i = c; i < JD * LS_F; i += 512)
      if (i % LS_F >= LS_NPRE) oxb[i] = 0.f;

  if (with_A) {          // exact-order fp32 path; the tensor-core path computes A with lsw_audio_proj
    float acc[LS_F];
#pragma unroll
    for (int f = 0; f < LS_F; ++f) acc[f] = 0.f;
    for (int k = 0; k < LS_AF; ++k) {
      const float wv = w.w_a_t[(size_t)k * LS_D + c];
#pragma unroll
      for (int f = 0; f < LS_F; ++f) acc[f] = fmaf(af_s[k * LS_F + f], wv, acc[f]);
    }
    float* Ab = A + (size_t)b * LS_F * LS_D;
#pragma unroll
    for (int f = 0; f < LS_F; ++f) Ab[f * LS_D + c] = acc[f];
  }

  float pre[LS_NPRE] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 9
  for (int j = 0; j < JD; ++j) {
    const float wv = w.w_o_t[(size_t)j * LS_D + c];
#pragma unroll
    for (int f = 0; f < LS_NPRE; ++f) pre[f] = fmaf(ox_s[j * LS_NPRE + f], wv, pre[f]);
  }
  const float bin = w.b_in[c], wbit = w.w_bit[c];
  float* Pb = P + (size_t)b * LS_F * LS_D;
#pragma unroll
  for (int f = 0; f < LS_F; ++f) Pb[f * LS_D + c] = (f < LS_NPRE) ? (pre[f] + wbit) + bin : bin;

  float mu = 0.f, lv = 0.f;
#pragma unroll 16
  for (int k = 0; k < LS_SPK; ++k) {          // 16 x 2 weight loads in flight (one pair at a time: 256 L2 round trips)
    const float zv = z_s[k];
    mu = fmaf(zv, w.w_mu_t[(size_t)k * LS_D + c], mu);
    lv = fmaf(zv, w.w_lv_t[(size_t)k * LS_D + c], lv);
  }
  z_mu[(size_t)b * LS_D + c] = mu + w.b_mu[c];
  z_lv[(size_t)b * LS_D + c] = lv + w.b_lv[c];
  if (w.emo_emb != nullptr && emo != nullptr) {
    long long e = emo[(size_t)b * emo_stride];
    e = e < 0 ? 0 : (e >= n_emo ? n_emo - 1 : e);
    emo_tok[(size_t)b * LS_D + c] = w.emo_emb[(size_t)e * LS_D + c];
  }
}

int lsk_cond_proj(ls_handle* h, int B, int b0, const float* af_cm, float* origin_x, const int64_t* vid,
                  const int64_t* emo, int64_t emo_stride, int mutate_origin, cudaStream_t s) {
  const size_t smem = (size_t)(LS_AF * LS_F + h->JD * LS_NPRE + LS_SPK) * sizeof(float);
  // the attribute is per DEVICE: set on every launch (one process may drive several GPUs)
  LS_CUDA(h, cudaFuncSetAttribute(cond_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const bool tc = lsw_available(h) && ls_get_impl(h) != LS_IMPL_SIMT;
  if (tc) {
    const int rc = lsw_audio_proj(h, af_cm, B, h->A + (size_t)b0 * LS_F * LS_D, s);
    if (rc) return rc;
  }
  cond_proj_kernel<<<B, 512, smem, s>>>(h->w, h->JD, h->cfg.n_speakers, h->cfg.n_emotions, af_cm, origin_x, vid, emo,
                                        emo_stride, mutate_origin, b0, tc ? 0 : 1, h->A, h->P, h->z_mu, h->z_lv, h->emo_tok);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

// emb[t] = W2 * silu(W1 * pe[t] + b1) + b2 for every t < n_t  (mlp_module.py:135-136).
// w1t / w2t are the transposed weights [k][c].
__global__ void __launch_bounds__(512) time_embed_kernel(const float* __restrict__ pe, const float* __restrict__ w1t,
                                                         const float* __restrict__ b1, const float* __restrict__ w2t,
                                                         const float* __restrict__ b2, float* __restrict__ table) {
  __shared__ float e[LS_D];
  __shared__ float g[LS_D];
  const int t = blockIdx.x, c = threadIdx.x;
  e[c] = pe[(size_t)t * LS_D + c];
  __syncthreads();
  float a = 0.f;
  for (int k = 0; k < LS_D; ++k) a = fmaf(e[k], w1t[(size_t)k * LS_D + c], a);
  a += b1[c];
  g[c] = a / (1.f + expf(-a));
  __syncthreads();
  float o = 0.f;
  for (int k = 0; k < LS_D; ++k) o = fmaf(g[k], w2t[(size_t)k * LS_D + c], o);
  table[(size_t)t * LS_D + c] = o + b2[c];
}

int lsk_time_embed_table(ls_handle* h, const float* pe, const float* w1t, const float* b1, const float* w2t,
                         const float* b2, float* table, int n_t, cudaStream_t s) {
  time_embed_kernel<<<n_t, 512, 0, s>>>(pe, w1t, b1, w2t, b2, table);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

// out[c][r] = in[r*ld_in + c]   (rows x cols -> cols x rows)
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols, int ld_in) {
  __shared__ float tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (r0 + i < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)(r0 + i) * ld_in + c];
  __syncthreads();
  int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (c0 + i < cols && r < rows) out[(size_t)(c0 + i) * rows + r] = tile[threadIdx.x][i];
}

int lsk_transpose(ls_handle* h, const float* in, float* out, int rows, int cols, int ld_in, cudaStream_t s) {
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, s>>>(in, out, rows, cols, ld_in);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}
