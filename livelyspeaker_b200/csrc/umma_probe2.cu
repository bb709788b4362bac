// Probe for the round-2 design (DESIGN.md 4.1 "next step"): one tcgen05.mma.cta_group::2 over a cluster pair.
// M = 256 (128 rows of A from EACH CTA's shared memory), N = 80 (40 rows of B from each CTA), K = 64, kind::f16,
// small-integer operands so the result is exact.  Each CTA's tensor memory receives ITS 128 rows of D for ALL N
// columns.  Hypothesis checked: CTA r supplies A rows [128r, 128r+128) and B rows [40r, 40r+40).
// Build: make umma_probe2 ; run on the B200 (wrap in `timeout`: a wrong guess about the pair protocol can hang).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "ls_tc.cuh"
using namespace lstc;

constexpr int N = 80, K = 64;

__device__ __host__ inline float wval(int m, int k) { return (float)((m * 2 + k * 3) % 5 - 2); }
__device__ __host__ inline float uval(int n, int k) { return (float)((n * 3 + k * 5) % 7 - 3); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2_kernel(float* __restrict__ d_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* A = sm;                    // 128 x 64 bf16, K-major SW128 (16 KB)
  uint8_t* B = sm + 16384;            // 40 x 64 bf16 (5 row groups)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 16384 + 8192 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = tid; i < 128 * K; i += 128) {
    const int m = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(A + tile_off(m, k, 0)) = __float2bfloat16(wval(128 * rank + m, k));
  }
  for (int i = tid; i < (N / 2) * K; i += 128) {
    const int n = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(B + tile_off(n, k, 0)) = __float2bfloat16(uval((N / 2) * rank + n, k));
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = *slot;
  if (rank == 0 && warp == 0) {
    constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
    const uint32_t idesc = idesc_bf16(256, N, 0, 0);
    const uint32_t al = desc_lo32(smem_u32(A), 16), bl = desc_lo32(smem_u32(B), 16);
    for (uint32_t ks = 0; ks < 4; ++ks) {
      const uint32_t acc = ks ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
          "mov.b64 da, {%1, %2};\n\t"
          "mov.b64 db, {%3, %4};\n\t"
          "setp.ne.b32 p, %6, 0;\n\t"
          "elect.sync _|q, 0xffffffff;\n\t"
          "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
          ::"r"(tmem), "r"(al + 2 * ks), "r"(DH), "r"(bl + 2 * ks), "r"(DH), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
  }
  mbar_wait(bar, 0);
  __syncwarp();
  tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) d_out[(size_t)(128 * rank + tid) * N + c0 + i] = v[i];
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 256 * N * 4);
  cudaMemset(d, 0xff, 256 * N * 4);
  const int smem = 16384 + 8192 + 64 + 1024;
  cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe2_kernel<<<2, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> D(256 * N);
  cudaMemcpy(D.data(), d, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad[2][2] = {{0, 0}, {0, 0}};
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)wval(m, k) * uval(n, k);
      if (!(fabs(ref - D[m * N + n]) <= 1e-3)) ++bad[m / 128][n / 40];
    }
  printf("cta_group::2 M=256 N=80 K=64: mismatches by (row half, column half): [%d %d] [%d %d]  D[0][0..3] = %g %g %g %g  D[128][40..43] = %g %g %g %g\n",
         bad[0][0], bad[0][1], bad[1][0], bad[1][1], D[0], D[1], D[2], D[3], D[128 * N + 40], D[128 * N + 41], D[128 * N + 42], D[128 * N + 43]);
  printf((bad[0][0] | bad[0][1] | bad[1][0] | bad[1][1]) == 0 ? "PROBE2 OK: CTA r supplies A rows [128r,+128) and B rows [N/2*r,+N/2); D rows split by CTA, all N columns in both\n" : "PROBE2 MISMATCH\n");
  return 0;
}
