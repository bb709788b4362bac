// Post-sampling rhythm metric of the TED evaluation on the device:
//   scripts/test_RAG_ted.py:84-111  direction vectors -> joint-pair angles -> per-frame angle change -> motion beats
//   scripts/test_RAG_ted.py:112-123 beat alignment score against audio onset times
// Stateless entry points (no ls_handle), fp32 in the reference's op order for the angles, fp64 for the score like the
// reference's numpy arithmetic.  One CTA per clip: the work is a few hundred flops per clip, the point of having it
// here is that the sampler's output never leaves the device between the loop and the metric.
#include <cstring>

#include "ls_internal.cuh"

namespace {

struct BeatParams {
  int njoints, n_frames, n_pairs;
  float thres;
  float mean[LS_METRIC_MAX_JOINTS * 3];
  int pair[LS_METRIC_MAX_PAIRS][2];
  float change[LS_METRIC_MAX_PAIRS];
};

constexpr int MAX_F = 64;

__global__ void __launch_bounds__(MAX_F, 8)
motion_beats_kernel(const BeatParams p, const float* __restrict__ sample, float* __restrict__ angle_diff,
                    uint8_t* __restrict__ beat_mask) {
  __shared__ float ang[LS_METRIC_MAX_PAIRS][MAX_F];
  __shared__ float diff[MAX_F];
  const int b = blockIdx.x, f = threadIdx.x, F = p.n_frames;
  const float* sb = sample + (size_t)b * p.njoints * 3 * F;
  if (f < F) {
    for (int q = 0; q < p.n_pairs; ++q) {
      float v[2][3];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int j = p.pair[q][s];
        // aligned_motions + mean_dir_vec (:88), F.normalize over the 3 components (:90): x / max(||x||, 1e-12)
        const float x = __fadd_rn(sb[(size_t)(j * 3 + 0) * F + f], p.mean[j * 3 + 0]);
        const float y = __fadd_rn(sb[(size_t)(j * 3 + 1) * F + f], p.mean[j * 3 + 1]);
        const float z = __fadd_rn(sb[(size_t)(j * 3 + 2) * F + f], p.mean[j * 3 + 2]);
        const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))), 1e-12f);
        v[s][0] = __fdiv_rn(x, nrm);
        v[s][1] = __fdiv_rn(y, nrm);
        v[s][2] = __fdiv_rn(z, nrm);
      }
      float ip = __fadd_rn(__fadd_rn(__fmul_rn(v[0][0], v[1][0]), __fmul_rn(v[0][1], v[1][1])), __fmul_rn(v[0][2], v[1][2]));
      ip = fminf(fmaxf(ip, -1.f), 1.f);                                     // :97
      ang[q][f] = __fdiv_rn(acosf(ip), 3.14159265358979323846f);           // :98 (math.pi as an fp32 scalar)
    }
  }
  __syncthreads();
  if (f < F) {
    float d = 0.f;                                                          // :104 the prepended zero column
    if (f >= 1) {
      for (int q = 0; q < p.n_pairs; ++q) {                                 // :100-103, pairs accumulated in order
        const float t = __fdiv_rn(__fdiv_rn(fabsf(__fsub_rn(ang[q][f], ang[q][f - 1])), p.change[q]), (float)p.n_pairs);
        d = (q == 0) ? t : __fadd_rn(d, t);
      }
    }
    diff[f] = d;
    angle_diff[(size_t)b * F + f] = d;
  }
  __syncthreads();
  if (f < F) {
    uint8_t m = 0;
    if (f >= 2 && f < F - 1) {                                              // :107 for t in range(2, 33)
      const float a = diff[f], lo = diff[f - 1], hi = diff[f + 1];
      if (a < lo && a < hi && (__fsub_rn(lo, a) >= p.thres || __fsub_rn(hi, a) >= p.thres)) m = 1;
    }
    beat_mask[(size_t)b * F + f] = m;
  }
}

__global__ void __launch_bounds__(128, 8)
beat_align_kernel(int F, int M, double fps, double sigma, const uint8_t* __restrict__ beat_mask,
                  const float* __restrict__ audio_beats, const int* __restrict__ n_audio, double* __restrict__ clip_score,
                  int* __restrict__ clip_n_motion, int* __restrict__ clip_n_audio) {
  __shared__ double part[128];
  __shared__ int n_motion;
  const int b = blockIdx.x, tid = threadIdx.x;
  const uint8_t* mb = beat_mask + (size_t)b * F;
  if (tid == 0) {
    int n = 0;
    for (int t = 0; t < F; ++t) n += mb[t] ? 1 : 0;
    n_motion = n;
  }
  __syncthreads();
  const int na = min(max(n_audio[b], 0), M);
  double s = 0.0;
  if (n_motion > 0) {
    for (int i = tid; i < na; i += 128) {
      const double ab = (double)audio_beats[(size_t)b * M + i];
      double best = 1e300;
      for (int t = 0; t < F; ++t)
        if (mb[t]) {
          const double d = ab - (double)t / fps;                            // :110 float(t) / 15.0
          best = fmin(best, d * d);
        }
      s += exp(-best / (2.0 * sigma * sigma));                              // :120
    }
  }
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int i = 0; i < 128; ++i) tot += part[i];
    clip_score[b] = tot;
    clip_n_motion[b] = n_motion;
    clip_n_audio[b] = n_motion > 0 ? na : 0;                                // :117-118 clips without motion beats are skipped
  }
}

}  // namespace

extern "C" int ls_motion_beats(int32_t B, int32_t njoints, int32_t n_frames, const float* sample,
                               const float* mean_dir_vec_host, const int32_t* angle_pairs_host,
                               const float* change_angle_host, int32_t n_pairs, float thres, float* angle_diff,
                               uint8_t* beat_mask, int32_t device, void* stream) {
  if (B < 1 || !sample || !mean_dir_vec_host || !angle_pairs_host || !change_angle_host || !angle_diff || !beat_mask)
    return ls_fail(nullptr, LS_EINVAL, "ls_motion_beats: bad argument");
  if (njoints < 1 || njoints > LS_METRIC_MAX_JOINTS || n_pairs < 1 || n_pairs > LS_METRIC_MAX_PAIRS || n_frames < 3 ||
      n_frames > MAX_F)
    return ls_fail(nullptr, LS_EUNSUPPORTED, "ls_motion_beats: <= %d joints, <= %d pairs, 3..%d frames", LS_METRIC_MAX_JOINTS,
                   LS_METRIC_MAX_PAIRS, MAX_F);
  BeatParams p{};
  p.njoints = njoints;
  p.n_frames = n_frames;
  p.n_pairs = n_pairs;
  p.thres = thres;
  memcpy(p.mean, mean_dir_vec_host, sizeof(float) * njoints * 3);
  for (int q = 0; q < n_pairs; ++q) {
    p.pair[q][0] = angle_pairs_host[2 * q];
    p.pair[q][1] = angle_pairs_host[2 * q + 1];
    if (p.pair[q][0] < 0 || p.pair[q][0] >= njoints || p.pair[q][1] < 0 || p.pair[q][1] >= njoints)
      return ls_fail(nullptr, LS_EINVAL, "ls_motion_beats: joint index out of range in pair %d", q);
    p.change[q] = change_angle_host[q];
  }
  if (cudaSetDevice(device) != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_motion_beats: cudaSetDevice(%d)", device);
  motion_beats_kernel<<<B, MAX_F, 0, (cudaStream_t)stream>>>(p, sample, angle_diff, beat_mask);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_motion_beats: %s", cudaGetErrorString(e));
  return LS_OK;
}

extern "C" int ls_beat_align(int32_t B, int32_t n_frames, const uint8_t* beat_mask, const float* audio_beats,
                             const int32_t* n_audio, int32_t M, float fps, float sigma, double* clip_score,
                             int32_t* clip_n_motion, int32_t* clip_n_audio, int32_t device, void* stream) {
  if (B < 1 || n_frames < 1 || M < 1 || !beat_mask || !audio_beats || !n_audio || !clip_score || !clip_n_motion ||
      !clip_n_audio || !(fps > 0.f) || !(sigma > 0.f))
    return ls_fail(nullptr, LS_EINVAL, "ls_beat_align: bad argument");
  if (cudaSetDevice(device) != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_beat_align: cudaSetDevice(%d)", device);
  beat_align_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(n_frames, M, (double)fps, (double)sigma, beat_mask, audio_beats,
                                                          n_audio, clip_score, clip_n_motion, clip_n_audio);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_beat_align: %s", cudaGetErrorString(e));
  return LS_OK;
}


// ---- training loss terms: GaussianDiffusion.training_losses, HUBER branch, forward values --------
// scripts/diffusion/gaussian_diffusion.py:21-24 (compute_huber = smooth_l1(a / 0.1, b / 0.1) * 0.1, mean over all
// elements), :1379-1391 (rot_mse on the sample, vel_mse on its frame differences, kld of the style posterior).  One
// block, fixed summation order (deterministic), fp64 accumulation.
namespace {
__device__ __forceinline__ double huber01(float a, float b) {
  const float d = fabsf(a / 0.1f - b / 0.1f);
  return d < 1.f ? 0.5 * (double)d * (double)d : (double)d - 0.5;
}
__global__ void __launch_bounds__(1024) huber_terms_kernel(long long rows, int F, const float* __restrict__ target,
                                                           const float* __restrict__ output, long long n_z,
                                                           const float* __restrict__ z_mu, const float* __restrict__ z_lv,
                                                           float* __restrict__ terms) {
  __shared__ double red[3][32];
  double s_rot = 0.0, s_vel = 0.0, s_kld = 0.0;
  for (long long r = threadIdx.x; r < rows; r += 1024) {
    const float* tr = target + r * F;
    const float* orow = output + r * F;
    float tp = tr[0], op = orow[0];
    s_rot += huber01(tp, op);
    for (int f = 1; f < F; ++f) {
      const float tc = tr[f], oc = orow[f];
      s_rot += huber01(tc, oc);
      s_vel += huber01(tc - tp, oc - op);
      tp = tc;
      op = oc;
    }
  }
  if (z_mu != nullptr)
    for (long long i = threadIdx.x; i < n_z; i += 1024) {
      const float mu = z_mu[i], lv = z_lv[i];
      s_kld += (double)(1.f + lv - mu * mu - expf(lv));
    }
  double v[3] = {s_rot, s_vel, s_kld};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[threadIdx.x][w];
    const double n = threadIdx.x == 0 ? (double)rows * F : threadIdx.x == 1 ? (double)rows * (F - 1) : (double)n_z;
    terms[threadIdx.x] = threadIdx.x == 2 ? (float)(-0.5 * t / n) : (float)(0.1 * t / n);
  }
}
}  // namespace

extern "C" int ls_huber_terms(int64_t rows, int32_t n_frames, const float* target, const float* output, int64_t n_z,
                              const float* z_mu, const float* z_logvar, float* terms, int32_t device, void* stream) {
  if (rows < 1 || n_frames < 2 || !target || !output || !terms || (z_mu != nullptr) != (z_logvar != nullptr))
    return ls_fail(nullptr, LS_EINVAL, "ls_huber_terms: bad argument");
  if (cudaSetDevice(device) != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "cudaSetDevice failed");
  huber_terms_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rows, n_frames, target, output, n_z, z_mu, z_logvar, terms);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_huber_terms: %s", cudaGetErrorString(e));
  return LS_OK;
}


// ---- variational-bound terms (GaussianDiffusion._vb_terms_bpd / _prior_bpd / calc_bpd_loop) --------------------------
// scripts/diffusion/gaussian_diffusion.py:1213-1247, 1573-1590 and scripts/diffusion/losses.py:12-77: per clip the mean
// over all elements of normal_kl(true posterior || model) or, at t == 0, of the discretised Gaussian negative
// log-likelihood of x_start, in bits.  With LivelySpeaker's fixed variances both log-variances are one number per clip.
// fp32 element arithmetic in the reference's operation order (no contraction), one block per clip, fixed-order fp64 sum.
namespace {
__device__ __forceinline__ float approx_normal_cdf(float x) {                  // losses.py:46-51
  const float x3 = __fmul_rn(__fmul_rn(x, x), x);                               // torch.pow(x, 3)
  const float inner = __fmul_rn(0.7978845608028654f, __fadd_rn(x, __fmul_rn(0.044715f, x3)));
  return __fmul_rn(0.5f, __fadd_rn(1.0f, tanhf(inner)));
}
__global__ void __launch_bounds__(256) vb_terms_kernel(long long n, const float* __restrict__ x_start,
                                                       const float* __restrict__ mean1, const float* __restrict__ mean2,
                                                       const float* __restrict__ logvar1, const float* __restrict__ logvar2,
                                                       const long long* __restrict__ t, float* __restrict__ out) {
  __shared__ double red[8];
  const int b = blockIdx.x;
  const float lv1 = logvar1[b], lv2 = logvar2 ? logvar2[b] : 0.f;
  const bool nll = t != nullptr && t[b] == 0;
  const float* xs = x_start + (size_t)b * n;
  const float* m1 = mean1 + (size_t)b * n;
  const float* m2 = mean2 ? mean2 + (size_t)b * n : nullptr;
  // normal_kl (losses.py:12-43): 0.5 * (-1 + lv2 - lv1 + exp(lv1 - lv2) + (m1 - m2)^2 * exp(-lv2))
  const float head = __fadd_rn(__fadd_rn(__fadd_rn(-1.0f, lv2), -lv1), expf(__fadd_rn(lv1, -lv2)));
  const float inv_var2 = expf(-lv2);
  const float inv_stdv = expf(-__fmul_rn(0.5f, lv2));                           // exp(-log_scales), log_scales = 0.5 * lv2
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const float mm = m2 ? m2[i] : 0.f;
    float v;
    if (!nll) {
      const float d = __fadd_rn(m1[i], -mm);
      v = __fmul_rn(0.5f, __fadd_rn(head, __fmul_rn(__fmul_rn(d, d), inv_var2)));
    } else {                                                                    // losses.py:54-77, negated
      const float x = xs[i], c = __fadd_rn(x, -mm);
      const float cdf_plus = approx_normal_cdf(__fmul_rn(inv_stdv, __fadd_rn(c, 1.0f / 255.0f)));
      const float cdf_min = approx_normal_cdf(__fmul_rn(inv_stdv, __fadd_rn(c, -(1.0f / 255.0f))));
      float lp;
      if (x < -0.999f) lp = logf(fmaxf(cdf_plus, 1e-12f));
      else if (x > 0.999f) lp = logf(fmaxf(__fadd_rn(1.0f, -cdf_min), 1e-12f));
      else lp = logf(fmaxf(__fadd_rn(cdf_plus, -cdf_min), 1e-12f));
      v = -lp;
    }
    s += (double)v;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    out[b] = (float)(tot / (double)n) / 0.6931471805599453f;                    // mean_flat(.) / np.log(2.0)
  }
}
}  // namespace

extern "C" int ls_vb_terms(int32_t B, int64_t n, const float* x_start, const float* mean1, const float* mean2,
                           const float* logvar1, const float* logvar2, const int64_t* t, float* out, int32_t device,
                           void* stream) {
  if (B < 1 || n < 1 || !mean1 || !logvar1 || !out || (t != nullptr && (!x_start || !mean2 || !logvar2)))
    return ls_fail(nullptr, LS_EINVAL, "ls_vb_terms: bad argument");
  if (cudaSetDevice(device) != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_vb_terms: cudaSetDevice(%d)", device);
  vb_terms_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((long long)n, x_start, mean1, mean2, logvar1, logvar2,
                                                        (const long long*)t, out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_vb_terms: %s", cudaGetErrorString(e));
  return LS_OK;
}
