// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SS operands)
// as a function of N and of the A-operand layout, operands resident in shared memory.
// Build: make umma_bench ; run on the B200.  Feeds DESIGN.md "what bounds the GEMM phases".
#include <cstdio>
#include <cstdlib>
#include "ls_tc.cuh"
using namespace lstc;

__global__ void __launch_bounds__(128, 1) bench_kernel(int N, int a_mn, int iters, int spread, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid * 4; i < 160 * 1024; i += 128 * 4) *reinterpret_cast<uint32_t*>(sm + i) = 0x3c003c00u;  // finite bf16
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&tslot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tslot;
  if (tid == 0) {
    const uint32_t a_s = smem_u32(sm), b_s = smem_u32(sm + 64 * 1024);
    const uint32_t idesc = idesc_bf16(128, N, a_mn, 0);
    constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
    const uint32_t al = a_mn ? desc_lo32(a_s, 9216) : desc_lo32(a_s, 16), bl = desc_lo32(b_s, 16);
    // warm-up
    for (int i = 0; i < 16; ++i) umma_bf16_split(tmem, al, DH, bl, DH, idesc, 1u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t ks = (i & 3);
      // spread != 0: every MMA reads a different A block (16 KB apart, 4 blocks) and B block
      const uint32_t blk = spread ? ((i >> 2) & 3) : 0;
      umma_bf16_split(tmem, al + (a_mn ? ks * 128 : ks * 2) + blk * 1024, DH, bl + ks * 2 + blk * 640, DH, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 170 * 1024;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2048;
  for (int spread = 0; spread < 2; ++spread)
  for (int a_mn = 0; a_mn < 2; ++a_mn)
    for (int N : {16, 40, 48, 64, 80, 96, 112, 128, 144, 160, 192, 256}) {
      bench_kernel<<<1, 128, smem>>>(N, a_mn, iters, spread, d);
      long long c;
      if (cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
      printf("%s A %s-major  N=%3d : %.1f cycles/MMA  (%.0f MAC/cycle)\n", spread ? "spread" : "reuse ", a_mn ? "MN" : "K ", N, (double)c / iters,
             128.0 * N * 16 * iters / c);
    }
  return 0;
}
