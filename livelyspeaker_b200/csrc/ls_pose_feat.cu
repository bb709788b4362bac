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
// One CTA encodes CLIPS = 4 clips from shared memory: 1.4 MFLOP and 0.67 MB of weights per clip, i.e. the weights are
// the traffic (read through L1 / L2 once per CTA, the same address across a warp for the convolutions, consecutive
// addresses for the transposed linear weights); 128 CTAs cover a 512-clip batch in one wave.  The sampler's output
// stays on the device between the loop and the metric.
#include "ls_internal.cuh"

namespace {

constexpr int CLIPS = 4;
constexpr int NT = 256;
constexpr int F0 = 34, F1 = 32, F2 = 30, F3 = 14, F4 = 12;     // frames after each convolution
constexpr int BUF_A = CLIPS * 64 * F2;                         // x (<= 56 x 34), conv2 out, conv4 out, fc2 out
constexpr int BUF_B = CLIPS * 32 * F1;                         // conv1 out, conv3 out, fc1 out, fc3 out

// out[g][co][p] = act(scale[co] * (bias[co] + sum_ci sum_k w[co][ci][k] * in[g][ci][p * S + k]) + shift[co])
template <int K, int S>
__device__ __forceinline__ void conv_layer(const float* __restrict__ in, float* __restrict__ out, int ci_n, int co_n, int l_in,
                                           int l_out, int n_clips, const float* __restrict__ w, const float* __restrict__ bias,
                                           const float* __restrict__ scale, const float* __restrict__ shift, float slope) {
  const int per_clip = co_n * l_out;
  for (int idx = threadIdx.x; idx < n_clips * per_clip; idx += NT) {
    const int g = idx / per_clip, r = idx - g * per_clip, co = r / l_out, p = r - co * l_out;
    const float* src = in + (size_t)g * ci_n * l_in + p * S;
    const float* wr = w + (size_t)co * ci_n * K;
    float acc = __ldg(bias + co);
    for (int ci = 0; ci < ci_n; ++ci) {
#pragma unroll
      for (int k = 0; k < K; ++k) acc = fmaf(__ldg(wr + ci * K + k), src[ci * l_in + k], acc);
    }
    if (scale) {
      acc = fmaf(acc, __ldg(scale + co), __ldg(shift + co));
      acc = acc > 0.f ? acc : slope * acc;
    }
    out[idx] = acc;
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
  conv_layer<3, 1>(bufA, bufB, dim, 32, F0, F1, CLIPS, w.c1_w, w.c1_b, w.c1_scale, w.c1_shift, w.slope_conv);
  __syncthreads();
  conv_layer<3, 1>(bufB, bufA, 32, 64, F1, F2, CLIPS, w.c2_w, w.c2_b, w.c2_scale, w.c2_shift, w.slope_conv);
  __syncthreads();
  conv_layer<4, 2>(bufA, bufB, 64, 64, F2, F3, CLIPS, w.c3_w, w.c3_b, w.c3_scale, w.c3_shift, w.slope_conv);
  __syncthreads();
  conv_layer<3, 1>(bufB, bufA, 64, 32, F3, F4, CLIPS, w.c4_w, w.c4_b, nullptr, nullptr, 0.f);   // [g][32][12] = flatten(1)
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
