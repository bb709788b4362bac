// Micro-benchmark 2: what sets the ~55-cycle floor of a cta_group::1 M=128 K=16 tcgen05.mma?
// Warp-uniform elect-issued MMAs (the form the fused kernel uses), fully unrolled, with
//   NACC = number of distinct accumulators written round-robin (dependent-accumulate chain or not),
//   N    = MMA N,
//   AMODE: 0 = same A block every MMA, 1 = A walks 4 blocks (as the weight ring does).
// Build: make umma_bench2 ; run on the B200.
#include <cstdio>
#include <cstdlib>
#include "ls_tc.cuh"
using namespace lstc;

template <int N, int NACC, int AMODE>
__global__ void __launch_bounds__(128, 1) bench_kernel(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid * 4; i < 160 * 1024; i += 128 * 4) *reinterpret_cast<uint32_t*>(sm + i) = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&tslot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tslot, 0);
  if (warp == 1) {
    const uint32_t a_s = smem_u32(sm), b_s = smem_u32(sm + 64 * 1024);
    constexpr uint32_t idesc = idesc_bf16(128, N, 0, 0);
    constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
    const uint32_t al = desc_lo32(a_s, 16), bl = desc_lo32(b_s, 16);
    constexpr uint32_t DSTRIDE = (NACC * N <= 512) ? N : 512 / NACC;
    for (int i = 0; i < 16; ++i) umma_bf16_split_elect(tmem, al, DH, bl, DH, idesc, 1u);
    umma_commit_s_elect(smem_u32(&bar));
    mbar_wait_s(smem_u32(&bar), 0);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t ks = j & 3, blk = AMODE ? ((j >> 2) & 3) : 0;
        umma_bf16_split_elect(tmem + (j % NACC) * DSTRIDE, al + ks * 2 + blk * 1024, DH, bl + ks * 2, DH, idesc, 1u);
      }
    }
    umma_commit_s_elect(smem_u32(&bar));
    mbar_wait_s(smem_u32(&bar), 1);
    long long t1 = clock64();
    if ((tid & 31) == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N, int NACC, int AMODE>
void run(long long* d) {
  const int smem = 170 * 1024, iters = 4096;
  cudaFuncSetAttribute(bench_kernel<N, NACC, AMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench_kernel<N, NACC, AMODE><<<1, 128, smem>>>(iters, d);
  long long c = 0;
  if (cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error N=%d\n", N); exit(1); }
  printf("N=%3d  accumulators=%d  A %s : %.1f cycles/MMA  (%.0f MAC/cycle)\n", N, NACC, AMODE ? "walks" : "fixed",
         (double)c / iters, 128.0 * N * 16 * iters / c);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run<16, 1, 0>(d);  run<16, 2, 0>(d);  run<16, 4, 0>(d);
  run<40, 1, 0>(d);  run<40, 2, 0>(d);  run<40, 4, 0>(d);  run<40, 4, 1>(d);
  run<48, 1, 0>(d);  run<48, 2, 0>(d);  run<48, 4, 1>(d);
  run<80, 1, 0>(d);  run<80, 1, 1>(d);  run<80, 2, 0>(d);  run<80, 2, 1>(d);  run<80, 4, 0>(d);  run<80, 4, 1>(d);
  run<96, 1, 0>(d);  run<96, 2, 1>(d);  run<96, 4, 1>(d);
  run<128, 1, 0>(d); run<128, 2, 1>(d);
  run<144, 1, 0>(d); run<144, 2, 1>(d); run<144, 3, 1>(d);
  run<160, 1, 1>(d); run<160, 3, 1>(d);
  run<256, 1, 1>(d); run<256, 2, 1>(d);
  return 0;
}
