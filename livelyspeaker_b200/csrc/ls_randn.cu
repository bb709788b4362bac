// torch.randn-compatible normal draws for many tensors in ONE launch.
//
// The reference draws three tensors per denoising step with torch.randn / randn_like (gaussian_diffusion.py:543,
// RAG.py:10-13); a fused chunk of 16 steps therefore needs 48 tiny torch kernels.  This kernel reproduces, tensor
// by tensor, exactly what ATen's CUDA normal_ does (ATen/native/cuda/DistributionTemplates.h:
// calc_execution_policy + distribution_elementwise_grid_stride_kernel + curand_normal4): Philox4x32-10 seeded with
// the generator's seed, subsequence = global thread index of a 256-thread grid of
// min(#SM * (maxThreadsPerSM / 256), ceil(numel / 256)) blocks, offset = the generator's Philox offset, which every
// tensor advances by ((numel - 1) / (256 * grid * 4) + 1) * 4.  Element li of a dense tensor is written at memory
// offset li (TensorIterator walks dense tensors in memory order, also permuted ones).  The Python side checks the
// result bit for bit against torch before it trusts it (livelyspeaker_b200/gaussian_diffusion.py::_FusedDraws) and
// moves the generator's offset by the total increment, so the draw stream stays the reference's.
#include <curand_kernel.h>

#include "ls_internal.cuh"

namespace {

constexpr int RN_BLOCK = 256, RN_UNROLL = 4, RN_MAX = LS_RANDN_MAX_TENSORS;

struct RandnJob {
  float* out;
  unsigned long long offset;
  unsigned int numel, grid, iters;     // iters = Philox calls per thread of torch's grid: (numel - 1) / (256 * grid * 4) + 1
};
struct RandnJobs {
  RandnJob j[RN_MAX];
};

// Philox4x32-10 written out (curand_philox4x32_x.h: ten rounds, the key bumped by the Weyl constants between them).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int lo0 = 0xD2511F53u * c.x, hi0 = __umulhi(0xD2511F53u, c.x);
    const unsigned int lo1 = 0xCD9E8D57u * c.z, hi1 = __umulhi(0xCD9E8D57u, c.z);
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// What curand_init(seed, subsequence = idx, offset) followed by the it-th curand_normal4 returns, without the generator
// state: curand_init leaves the counter at (offset / 4, idx) - torch's offsets are multiples of 4 - and every curand4
// call returns Philox(counter) and advances it by one.  Going through curandStatePhilox4_32_10_t costs two Philox
// evaluations per draw (curand_init and curand4 each compute one ahead) plus the state traffic: 156 -> 101 us per chunk
// (ncu), and 76 us with 32-bit index arithmetic and the trip count computed on the host.  What is left is torch's mapping itself: at these sizes a thread of ATen's grid uses 1 or 2 of the 4 values of its
// Philox call, so 14 M Philox evaluations feed 16 M normals; a persistent grid walking the (tensor, block) pairs instead of
// 57k short blocks was tried and is not faster (110 us): the kernel is bound by its integer instruction stream.
// Only the Box-Muller pairs that land inside the tensor are evaluated (_curand_box_muller, curand_normal.h:70-90, the
// function curand_normal4 is made of).
__global__ void __launch_bounds__(RN_BLOCK) randn_torch_compat_kernel(const __grid_constant__ RandnJobs jobs,
                                                                      unsigned long long seed) {
  const RandnJob& jb = jobs.j[blockIdx.y];
  if (blockIdx.x >= jb.grid) return;
  // 32-bit index arithmetic (numel < 2^31, so every position below numel + 4 * stride fits in 32 bits) and the trip
  // count from the host: a 64-bit division per thread cost as much as the Philox evaluation itself.
  const unsigned int idx = blockIdx.x * RN_BLOCK + threadIdx.x, stride = RN_BLOCK * jb.grid, numel = jb.numel;
  const uint2 key = make_uint2((unsigned int)seed, (unsigned int)(seed >> 32));
  unsigned long long ctr = jb.offset >> 2;
  float* __restrict__ out = jb.out;
  unsigned int li = idx;
  for (unsigned int it = 0; it < jb.iters; ++it, ++ctr, li += 4u * stride) {
    if (li >= numel) break;              // later iterations lie further out still
    const uint4 v = philox4x32_10(make_uint4((unsigned int)ctr, (unsigned int)(ctr >> 32), idx, 0u), key);
    const float2 a = _curand_box_muller(v.x, v.y);
    out[li] = a.x;
    if (li + stride < numel) out[li + stride] = a.y;
    if (li + 2u * stride < numel) {
      const float2 b = _curand_box_muller(v.z, v.w);
      out[li + 2u * stride] = b.x;
      if (li + 3u * stride < numel) out[li + 3u * stride] = b.y;
    }
  }
}

}  // namespace

extern "C" int ls_randn_torch_compat(int32_t n, float* const* outs, const int64_t* numels, uint64_t seed,
                                     uint64_t philox_offset, uint64_t* total_increment, int32_t device, void* stream) {
  if (n < 1 || n > RN_MAX || !outs || !numels || !total_increment)
    return ls_fail(nullptr, LS_EINVAL, "ls_randn_torch_compat: 1..%d tensors", RN_MAX);
  if (philox_offset & 3ull)
    return ls_fail(nullptr, LS_EINVAL, "ls_randn_torch_compat: the Philox offset must be a multiple of 4 (torch's CUDA generator keeps it so)");
  int sms = 0, tpsm = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&tpsm, cudaDevAttrMaxThreadsPerMultiProcessor, device) != cudaSuccess)
    return ls_fail(nullptr, LS_ECUDA, "ls_randn_torch_compat: cannot query device %d", device);
  const unsigned int grid_cap = (unsigned int)sms * (unsigned int)(tpsm / RN_BLOCK);
  RandnJobs jobs{};
  unsigned long long off = philox_offset;
  unsigned int max_grid = 1;
  for (int i = 0; i < n; ++i) {
    if (!outs[i] || numels[i] < 1 || numels[i] > 0x7fffffffLL)
      return ls_fail(nullptr, LS_EINVAL, "ls_randn_torch_compat: tensor %d", i);
    const unsigned long long numel = (unsigned long long)numels[i];
    unsigned int grid = (unsigned int)((numel + RN_BLOCK - 1) / RN_BLOCK);
    if (grid > grid_cap) grid = grid_cap;
    const unsigned long long iters = (numel - 1) / ((unsigned long long)RN_BLOCK * grid * RN_UNROLL) + 1;
    jobs.j[i] = RandnJob{outs[i], off, (unsigned int)numel, grid, (unsigned int)iters};
    off += iters * 4ull;
    if (grid > max_grid) max_grid = grid;
  }
  for (int i = n; i < RN_MAX; ++i) jobs.j[i] = RandnJob{nullptr, 0, 0, 0, 0};
  randn_torch_compat_kernel<<<dim3(max_grid, n), RN_BLOCK, 0, (cudaStream_t)stream>>>(jobs, seed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "ls_randn_torch_compat: %s", cudaGetErrorString(e));
  *total_increment = off - philox_offset;
  return LS_OK;
}
