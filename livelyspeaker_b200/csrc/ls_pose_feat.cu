// Latent features of the TED evaluation's pose autoencoder ("FGD feature extraction"):
//   scripts/model/embedding_net.py:40-79  PoseEncoderConv.forward (eval mode): poses [B, 34, dim] -> transpose ->
//     Conv1d(dim,32,3) BN LReLU(0.2) -> Conv1d(32,64,3) BN LReLU(0.2) -> Conv1d(64,64,4,stride 2) BN LReLU(0.2) ->
//     Conv1d(64,32,3) -> flatten [32 x 12] -> Linear(384,256) BN LReLU -> Linear(256,128) BN LReLU -> Linear(128,32) ->
//     fc_mu / fc_logvar (32 -> 32 each)
//   scripts/model/ted_evaluator.py:35-41   push_samples keeps the mu head of both the generated and the real clips
// BatchNorm runs on its running statistics (the evaluator calls .eval()); the host folds it to one scale / shift pair
// per channel.  The two LeakyReLU slopes are arguments: the reference writes nn.LeakyReLU(True) in the linear stack,
// which makes the slope 1.0 there (an identity), and the host passes whatever the loaded modules hold.
//
// One CTA encodes CLIPS = 4 clips out of two ping-pong activation buffers in shared memory; 128 CTAs cover a 512-clip
// batch in one wave.  1.4 MFLOP per clip on CUDA cores: the convolutions are register-tiled (4 output channels x 4 clips
// per thread) because one output per thread is bound by shared-memory reads (one LDS per FMA at 128 B/clk per SM), the
// weights come through L1 / L2 (the same address across a warp in the convolutions, consecutive addresses in the
// transposed linear weights).  Measured: 46 us per wave of 148 CTAs (B = 4096: 323 us, 18 TFLOP/s fp32); a push_samples
// of 2 x 512 clips costs less than 0.2 ms next to the 0.7 s sampling loop whose output it consumes on the device.
#include "ls_internal.cuh"

namespace {

constexpr int CLIPS = 4;
constexpr int NT = 256;
constexpr int F0 = 34, F1 = 32, F2 = 30, F3 = 14, F4 = 12;     // frames after each convolution
constexpr int BUF_A = CLIPS * 64 * F2;                         // x (<= 56 x 34), conv2 out, conv4 out, fc2 out
constexpr int BUF_B = CLIPS * 32 * F1;                         // conv1 out, conv3 out, fc1 out, fc3 out

// out[g][co][p] = act(scale[co] * (bias[co] + sum_ci sum_k w[co][ci][k] * in[g][ci][p * S + k]) + shift[co])
// A thread owns position p of CO_T consecutive output channels of ALL clips: one shared-memory read feeds CO_T
// accumulators and one weight read CLIPS of them (with one output per thread the layer is bound by the 128 B/clk of
// shared-memory reads: one LDS per FMA).
template <int K, int S, int CO_T>
__device__ __forceinline__ void conv_layer(const float* __restrict__ in, float* __restrict__ out, int ci_n, int co_n, int l_in,
                                           int l_out, const float* __restrict__ w, const float* __restrict__ bias,
                                           const float* __restrict__ scale, const float* __restrict__ shift, float slope) {
  const int per_clip = co_n * l_out, in_clip = ci_n * l_in, items = (co_n / CO_T) * l_out;
  for (int r = threadIdx.x; r < items; r += NT) {
    const int cg = r / l_out, p = r - cg * l_out, co0 = cg * CO_T;
    const float* src = in + p * S;
    const float* wr = w + (size_t)co0 * ci_n * K;
    float acc[CO_T][CLIPS];
#pragma unroll
    for (int t = 0; t < CO_T; ++t) {
      const float b = __ldg(bias + co0 + t);
#pragma unroll
      for (int g = 0; g < CLIPS; ++g) acc[t][g] = b;
    }
#pragma unroll 2
    for (int ci = 0; ci < ci_n; ++ci) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float xv[CLIPS], wv[CO_T];
#pragma unroll
        for (int g = 0; g < CLIPS; ++g) xv[g] = src[g * in_clip + ci * l_in + k];
#pragma unroll
        for (int t = 0; t < CO_T; ++t) wv[t] = __ldg(wr + (t * ci_n + ci) * K + k);
#pragma unroll
        for (int t = 0; t < CO_T; ++t)
#pragma unroll
          for (int g = 0; g < CLIPS; ++g) acc[t][g] = fmaf(wv[t], xv[g], acc[t][g]);
      }
    }
#pragma unroll
    for (int t = 0; t < CO_T; ++t) {
      const float sc = scale ? __ldg(scale + co0 + t) : 1.f, sh = scale ? __ldg(shift + co0 + t) : 0.f;
#pragma unroll
      for (int g = 0; g < CLIPS; ++g) {
        float v = acc[t][g];
        if (scale) {
          v = fmaf(v, sc, sh);
          v = v > 0.f ? v : slope * v;
        }
        out[g * per_clip + (co0 + t) * l_out + p] = v;
      }
    }
  }
}

// out[g][j] = act(scale[j] * (bias[j] + sum_k in[g][k] * wt[k][j]) + shift[j]); wt is the transposed weight [n_in][n_out]
__device__ __forceinline__ void linear_layer(const float* __restrict__ in, float* __restrict__ out, int n_in, int n_out,
                                             const float* __restrict__ wt, const float* __restrict__ bias,
                                             const float* __restrict__ scale, const float* __restrict__ shift, float slope) {
  for (int j = threadIdx.x; j < n_out; j += NT) {
    float acc[CLIPS];
    const float b = __ldg(bias + j);
#pragma unroll
    for (int g = 0; g < CLIPS; ++g) acc[g] = b;
#pragma unroll 8
    for (int k = 0; k < n_in; ++k) {
      const float wv = __ldg(wt + (size_t)k * n_out + j);
#pragma unroll
      for (int g = 0; g < CLIPS; ++g) acc[g] = fmaf(in[g * n_in + k], wv, acc[g]);
    }
#pragma unroll
    for (int g = 0; g < CLIPS; ++g) {
      float v = acc[g];
      if (scale) {
        v = fmaf(v, __ldg(scale + j), __ldg(shift + j));
        v = v > 0.f ? v : slope * v;
      }
      out[g * n_out + j] = v;
    }
  }
}

__global__ void __launch_bounds__(NT)
pose_features_kernel(const ls_pose_encoder_weights w, int B, const float* __restrict__ poses, float* __restrict__ mu,
                     float* __restrict__ logvar) {
  __shared__ float bufA[BUF_A];
  __shared__ float bufB[BUF_B];
  const int b0 = blockIdx.x * CLIPS, n = min(CLIPS, B - b0), dim = w.pose_dim;
  // poses [B][34][dim] -> channel-major x[g][dim][34] (embedding_net.py:66); clips past the batch are zero-filled so
  // the linear layers can run all CLIPS slots
  for (int i = threadIdx.x; i < CLIPS * F0 * dim; i += NT) {
    const int g = i / (F0 * dim), r = i - g * (F0 * dim), f = r / dim, d = r - f * dim;
    bufA[(g * dim + d) * F0 + f] = g < n ? __ldg(poses + (size_t)(b0 + g) * F0 * dim + r) : 0.f;
  }
  __syncthreads();
  conv_layer<3, 1, 4>(bufA, bufB, dim, 32, F0, F1, w.c1_w, w.c1_b, w.c1_scale, w.c1_shift, w.slope_conv);
  __syncthreads();
  conv_layer<3, 1, 4>(bufB, bufA, 32, 64, F1, F2, w.c2_w, w.c2_b, w.c2_scale, w.c2_shift, w.slope_conv);
  __syncthreads();
  conv_layer<4, 2, 4>(bufA, bufB, 64, 64, F2, F3, w.c3_w, w.c3_b, w.c3_scale, w.c3_shift, w.slope_conv);
  __syncthreads();
  conv_layer<3, 1, 2>(bufB, bufA, 64, 32, F3, F4, w.c4_w, w.c4_b, nullptr, nullptr, 0.f);   // [g][32][12] = flatten(1)
  __syncthreads();
  linear_layer(bufA, bufB, 32 * F4, 256, w.f1_wt, w.f1_b, w.f1_scale, w.f1_shift, w.slope_fc);
  __syncthreads();
  linear_layer(bufB, bufA, 256, 128, w.f2_wt, w.f2_b, w.f2_scale, w.f2_shift, w.slope_fc);
  __syncthreads();
  linear_layer(bufA, bufB, 128, 32, w.f3_wt, w.f3_b, nullptr, nullptr, 0.f);
  __syncthreads();
  // the two heads: thread = (clip, head, output)
  if (threadIdx.x < CLIPS * 64) {
    const int g = threadIdx.x >> 6, head = (threadIdx.x >> 5) & 1, j = threadIdx.x & 31;
    const float* wt = head ? w.lv_wt : w.mu_wt;
    float acc = __ldg((head ? w.lv_b : w.mu_b) + j);
    for (int k = 0; k < 32; ++k) acc = fmaf(bufB[g * 32 + k], __ldg(wt + k * 32 + j), acc);
    float* dst = head ? logvar : mu;
    if (g < n && dst) dst[(size_t)(b0 + g) * 32 + j] = acc;
  }
}

}  // namespace

extern "C" int ls_pose_features(const ls_pose_encoder_weights* w, int32_t B, const float* poses, float* mu, float* logvar,
                                int32_t device, void* stream) {
  if (!w || B < 1 || !poses || !mu) return ls_fail(nullptr, LS_EINVAL, "ls_pose_features: bad argument");
  if (w->n_frames != F0 || w->pose_dim < 1 || w->pose_dim * F0 * CLIPS > BUF_A)
    return ls_fail(nullptr, LS_EUNSUPPORTED, "ls_pose_features: n_frames must be %d (the 384-wide linear layer) and pose_dim <= %d",
                   F0, BUF_A / (F0 * CLIPS));
  const void* need[] = {w->c1_w, w->c1_b, w->c1_scale, w->c1_shift, w->c2_w, w->c2_b, w->c2_scale, w->c2_shift, w->c3_w, w->c3_b,
                        w->c3_scale, w->c3_shift, w->c4_w, w->c4_b, w->f1_wt, w->f1_b, w->f1_scale, w->f1_shift, w->f2_wt, w->f2_b,
                        w->f2_scale, w->f2_shift, w->f3_wt, w->f3_b, w->mu_wt, w->mu_b, w->lv_wt, w->lv_b};
  for (const void* p : need)
    if (!p) return ls_fail(nullptr, LS_EINVAL, "ls_pose_features: a weight pointer is null");
  if (cudaSetDevice(device) != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_pose_features: cudaSetDevice(%d)", device);
  pose_features_kernel<<<(B + CLIPS - 1) / CLIPS, NT, 0, (cudaStream_t)stream>>>(*w, B, poses, mu, logvar);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_pose_features: %s", cudaGetErrorString(e));
  return LS_OK;
}
