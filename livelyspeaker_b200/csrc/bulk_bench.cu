// Micro-benchmark: sustained global(L2) -> shared bandwidth of cp.async.bulk per SM, as a function
// of copy size and copies in flight, on 1 CTA and on all SMs.  Feeds DESIGN.md (weight ring sizing).
#include <cstdio>
#include <cstdlib>
#include "ls_tc.cuh"
using namespace lstc;

__global__ void __launch_bounds__(128, 1) bulk_kernel(const uint8_t* src, size_t src_bytes, int copy_bytes, int depth,
                                                      int n_copies, int n_issuers, int lanes_mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars_all[64];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(&bars_all[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // issuer k = lane 0 of warp k (lanes_mode 0) or lane k of warp 0 (lanes_mode 1)
  const int me = lanes_mode ? ((threadIdx.x < 32) ? (int)threadIdx.x : 99) : ((threadIdx.x & 31) == 0 ? (int)(threadIdx.x >> 5) : 99);
  if (me < n_issuers) {
    uint64_t* bars = bars_all + me * 8;
    sm += (size_t)me * depth * copy_bytes;
    const size_t base = ((size_t)(blockIdx.x * 4 + me) * 7919 * 16384) % (src_bytes - (size_t)copy_bytes * 4);
    long long t0 = clock64();
    for (int i = 0; i < n_copies + depth; ++i) {
      const int slot = i % depth;
      if (i >= depth) mbar_wait(&bars[slot], ((i / depth) - 1) & 1);      // previous copy in this slot landed
      if (i < n_copies) {
        mbar_arrive_expect_tx(&bars[slot], copy_bytes);
        size_t off = (base + (size_t)i * copy_bytes) % (src_bytes - copy_bytes);
        off &= ~size_t(15);
        bulk_g2s(sm + (size_t)slot * copy_bytes, src + off, copy_bytes, &bars[slot]);
      }
    }
    long long t1 = clock64();
    if (me == 0) out[blockIdx.x] = t1 - t0;
  }
}

int main() {
  const size_t src_bytes = 64u << 20;   // 64 MiB: L2 resident after the first pass
  uint8_t* src;
  long long* d;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  cudaMalloc(&d, 148 * 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int lanes_mode : {0, 1})
  for (int grid : {1, 148})
    for (int copy_bytes : {8192, 16384})
      for (int issuers : {1, 2, 4})
      for (int depth : {2, 4}) {
        if ((size_t)copy_bytes * depth * issuers > 196608) continue;
        const int n = 512;
        bulk_kernel<<<grid, 128, smem>>>(src, src_bytes, copy_bytes, depth, n, issuers, lanes_mode, d);   // warm L2
        bulk_kernel<<<grid, 128, smem>>>(src, src_bytes, copy_bytes, depth, n, issuers, lanes_mode, d);
        long long c[148];
        if (cudaMemcpy(c, d, grid * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = c[i] > mx ? c[i] : mx;
        printf("%s grid %3d  copy %5d B  issuers %d  in flight/issuer %d : %.1f B/cycle/SM  (%.0f cycles per copy per issuer)\n",
               lanes_mode ? "lanes" : "warps", grid, copy_bytes, issuers, depth, (double)copy_bytes * n * issuers / mx, (double)mx / n);
      }
  return 0;
}
