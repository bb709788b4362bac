// Standalone probe for the tcgen05 building blocks of the fused step kernel.  Not part of
// libls_b200.so; build with `make umma_probe` and run on the B200.  Every case computes a
// small-integer GEMM whose result is exact in fp32 and compares with the host.
//
//   case 0: A K-major (weights image, bulk-copied from global), B K-major (thread-written
//           operand tile), M=128 N=80 K=64, D at TMEM column 0
//   case 1: same, D at TMEM column 80 (unaligned accumulator base)
//   case 2: A MN-major view of the operand tile (M = channels, K = rows), B K-major
//           (token-mix weights), M=128 N=80 K=80, LBO = channel-block stride, SBO = 1024
//   case 3: as 2 with LBO / SBO swapped (to learn which reading of the ISA is right)
//   case 4: case 0 with K split over two 64-channel blocks + accumulate flag (K=128)
//   case 6: A operand from TENSOR MEMORY (TS mode, for a token-mix GEMM that keeps LayerNorm-1's output out of
//           shared memory): lane = M row, 32-bit column j = (k = 2j in the low half, k = 2j+1 in the high half),
//           8 columns per K=16 step, written with tcgen05.st.32x32b; B K-major from shared memory; M=128 N=80 K=64
//   case 7: as 6 with the halves swapped (to learn the packing if 6 fails)
//   case 5: the concatenated operand of the fused kernel: B = 144 contiguous rows (lo image rows
//           0..71 then hi image rows 0..71), one N=144 MMA per K step into D at column 152 (8- but not
//           16-aligned), then N=80 MMAs from row 72 accumulating into D + 72 (column 224)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#include "ls_tc.cuh"
using namespace lstc;

constexpr int ROWS = 80;                       // tile rows (tokens)
constexpr int RG = ROWS / 8;                   // 8-row groups
constexpr uint32_t CB_STRIDE = RG * 1024;      // bytes between 64-channel blocks of the operand tile
constexpr int NCH = 128;                       // channels in the probe tile (2 blocks)

struct Smem {
  alignas(1024) uint8_t u[2 * CB_STRIDE];      // operand tile: 80 rows x 128 channels
  alignas(1024) uint8_t w[2 * 16384];          // weight image(s): 128 x 64 bf16 each, K-major SW128
  alignas(1024) uint8_t wt[CB_STRIDE];         // token-mix weights: 80 x 80 (padded to 80 x 128?) K-major
  alignas(1024) uint8_t wt2[CB_STRIDE];
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(int which, const __nv_bfloat16* __restrict__ u_g /*[80][128]*/, const uint8_t* __restrict__ w_img,
             const __nv_bfloat16* __restrict__ wt_g /*[80][80]*/, float* __restrict__ d_out /*[128][80]*/) {
  extern __shared__ uint8_t raw[];
  Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&s.bar_w, 1);
    mbar_init(&s.bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&s.tmem_base);
  // operand tile written by threads in the swizzled layout
  for (int i = tid; i < ROWS * NCH; i += 128) {
    int r = i / NCH, c = i % NCH;
    *reinterpret_cast<__nv_bfloat16*>(s.u + tile_off(r, c, CB_STRIDE)) = u_g[i];
  }
  // token-mix weights [out=80][in=80] as a K-major tile: rows = out (N), K = in padded to 128
  for (int i = tid; i < ROWS * 128; i += 128) {
    int n = i / 128, k = i % 128;
    __nv_bfloat16 v = (k < ROWS) ? wt_g[n * ROWS + k] : __float2bfloat16(0.f);
    uint8_t* dst = (k < 64) ? s.wt : s.wt2;
    *reinterpret_cast<__nv_bfloat16*>(dst + tile_off(n, k & 63, 0)) = v;
  }
  if (which == 6 || which == 7) {     // A[m][k] = W[m][k], k < 64, packed bf16 pairs into TMEM columns 256..287 of lane m
    __syncthreads();
    const uint32_t tm = s.tmem_base;
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) {
      const float a0 = (float)((tid * 2 + (2 * j) * 3) % 5 - 2), a1 = (float)((tid * 2 + (2 * j + 1) * 3) % 5 - 2);
      const uint32_t b0 = __float_as_uint(a0) >> 16, b1 = __float_as_uint(a1) >> 16;      // exact bf16 of small integers
      r[j] = (which == 6) ? (b0 | (b1 << 16)) : (b1 | (b0 << 16));
    }
    for (int c0 = 0; c0 < 32; c0 += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
                       tm + ((uint32_t)(warp * 32) << 16) + 256 + c0),
                   "r"(r[c0]), "r"(r[c0 + 1]), "r"(r[c0 + 2]), "r"(r[c0 + 3]), "r"(r[c0 + 4]), "r"(r[c0 + 5]), "r"(r[c0 + 6]),
                   "r"(r[c0 + 7])
                   : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if (which == 5) {       // 160 rows x 64 channels, one block: row r = u_g row (r % 80), channel c -> c + 64*(r / 80)
    __syncthreads();
    for (int i = tid; i < 160 * 64; i += 128) {
      int r = i / 64, c = i % 64;
      *reinterpret_cast<__nv_bfloat16*>(s.u + tile_off(r, c, 0)) = u_g[(r % 80) * NCH + c + 64 * (r / 80)];
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  const uint32_t dcol = (which == 1) ? 80u : (which == 5) ? 152u : 0u;

  if (tid == 0) {
    if (which == 0 || which == 1 || which == 4) {
      const uint32_t nblk = (which == 4) ? 2 : 1;
      mbar_arrive_expect_tx(&s.bar_w, 16384 * nblk);
      for (uint32_t b = 0; b < nblk; ++b) bulk_g2s(s.w + b * 16384, w_img + b * 16384, 16384, &s.bar_w);
      mbar_wait(&s.bar_w, 0);
      tc_fence_after_sync();
      const uint32_t id = idesc_bf16(128, 80, 0, 0);
      for (uint32_t b = 0; b < nblk; ++b)
        for (uint32_t ks = 0; ks < 4; ++ks) {
          uint64_t ad = smem_desc(smem_u32(s.w) + b * 16384 + ks * 32, 16, 1024, SWZ_128B);
          uint64_t bd = smem_desc(smem_u32(s.u) + b * CB_STRIDE + ks * 32, 16, 1024, SWZ_128B);
          umma_bf16(tmem + dcol, ad, bd, id, (b | ks) ? 1u : 0u);
        }
    } else if (which == 6 || which == 7) {
      const uint32_t id = idesc_bf16(128, 80, 0, 0);
      for (uint32_t ks = 0; ks < 4; ++ks) {
        uint64_t bd = smem_desc(smem_u32(s.u) + ks * 32, 16, 1024, SWZ_128B);
        const uint32_t a_t = tmem + 256 + ks * 8, acc = ks ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
            ::"r"(tmem + dcol), "r"(a_t), "l"(bd), "r"(id), "r"(acc)
            : "memory");
      }
    } else if (which == 5) {
      mbar_arrive_expect_tx(&s.bar_w, 16384);
      bulk_g2s(s.w, w_img, 16384, &s.bar_w);
      mbar_wait(&s.bar_w, 0);
      tc_fence_after_sync();
      for (uint32_t ks = 0; ks < 4; ++ks) {
        uint64_t ad = smem_desc(smem_u32(s.w) + ks * 32, 16, 1024, SWZ_128B);
        uint64_t bd = smem_desc(smem_u32(s.u) + ks * 32, 16, 1024, SWZ_128B);
        umma_bf16(tmem + dcol, ad, bd, idesc_bf16(128, 144, 0, 0), ks ? 1u : 0u);
      }
      for (uint32_t ks = 0; ks < 4; ++ks) {
        uint64_t ad = smem_desc(smem_u32(s.w) + ks * 32, 16, 1024, SWZ_128B);
        uint64_t bd = smem_desc(smem_u32(s.u) + 9 * 1024 + ks * 32, 16, 1024, SWZ_128B);
        umma_bf16(tmem + dcol + 72, ad, bd, idesc_bf16(128, 80, 0, 0), 1u);
      }
    } else {
      // token mix: D[ch][tok_out] = sum_tok_in U[tok_in][ch] * Wt[tok_out][tok_in]
      const uint32_t id = idesc_bf16(128, 80, 1, 0);
      for (uint32_t ks = 0; ks < 5; ++ks) {       // 16 tokens (2 row groups) per step
        uint32_t lbo = (which == 2) ? CB_STRIDE : 1024u, sbo = (which == 2) ? 1024u : CB_STRIDE;
        uint64_t ad = smem_desc(smem_u32(s.u) + ks * 2048, lbo, sbo, SWZ_128B);
        // B: K-major, K = tok_in: 16 per step = 32 B inside the 64-wide block; steps 4.. are in wt2
        uint32_t bbase = (ks < 4) ? smem_u32(s.wt) + ks * 32 : smem_u32(s.wt2) + (ks - 4) * 32;
        uint64_t bd = smem_desc(bbase, 16, 1024, SWZ_128B);
        umma_bf16(tmem + dcol, ad, bd, id, ks ? 1u : 0u);
      }
    }
    umma_commit(&s.bar_mma);
  }
  mbar_wait(&s.bar_mma, 0);
  tc_fence_after_sync();
  // read back: warp w owns TMEM lanes 32w..32w+31 (= D rows)
  if (which == 5) {
    for (int c0 = 0; c0 < 144; c0 += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + dcol + c0, v);
      for (int i = 0; i < 8; ++i) d_out[(size_t)tid * 144 + c0 + i] = v[i];
    }
  } else
  for (int c0 = 0; c0 < 80; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + dcol + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) d_out[(size_t)tid * 80 + c0 + i] = v[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);   // exact for the small integers used here
}

int main() {
  std::vector<float> U(ROWS * NCH), W(128 * 128), WT(ROWS * ROWS);
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < NCH; ++c) U[r * NCH + c] = (float)((r * 3 + c * 5) % 7 - 3);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 128; ++k) W[m * 128 + k] = (float)((m * 2 + k * 3) % 5 - 2);
  for (int o = 0; o < ROWS; ++o)
    for (int i = 0; i < ROWS; ++i) WT[o * ROWS + i] = (float)((o + 2 * i) % 3 - 1);
  std::vector<uint16_t> Ub(U.size()), WTb(WT.size());
  for (size_t i = 0; i < U.size(); ++i) Ub[i] = f2bf(U[i]);
  for (size_t i = 0; i < WT.size(); ++i) WTb[i] = f2bf(WT[i]);
  // weight images: block b holds W[:, 64b:64b+64] as a K-major SW128 tile of 128 rows
  std::vector<uint8_t> img(2 * 16384);
  for (int b = 0; b < 2; ++b)
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 64; ++k) {
        uint16_t v = f2bf(W[m * 128 + 64 * b + k]);
        memcpy(&img[b * 16384 + tile_off(m, k, 0)], &v, 2);
      }
  void *dU, *dW, *dWT;
  float* dD;
  cudaMalloc(&dU, Ub.size() * 2);
  cudaMalloc(&dW, img.size());
  cudaMalloc(&dWT, WTb.size() * 2);
  cudaMalloc(&dD, 128 * 144 * 4);
  cudaMemcpy(dU, Ub.data(), Ub.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dWT, WTb.data(), WTb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int bad_total = 0;
  for (int which = 0; which < 8; ++which) {
    if (which == 5) continue;     // checked separately below
    cudaMemset(dD, 0xff, 128 * 80 * 4);
    probe_kernel<<<1, 128, smem>>>(which, (const __nv_bfloat16*)dU, (const uint8_t*)dW, (const __nv_bfloat16*)dWT, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("case %d: CUDA error %s\n", which, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> D(128 * 80);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 80; ++n) {
        double ref = 0;
        if (which == 0 || which == 1 || which == 4 || which >= 6) {
          int K = (which == 4) ? 128 : 64;
          for (int k = 0; k < K; ++k) ref += (double)W[m * 128 + k] * U[n * NCH + k];
        } else {
          for (int i = 0; i < ROWS; ++i) ref += (double)U[i * NCH + m] * WT[n * ROWS + i];
        }
        double err = fabs(ref - D[m * 80 + n]);
        if (!(err <= 1e-3)) ++bad;
        if (err > maxerr || err != err) maxerr = err;
      }
    printf("case %d: mismatches %d / %d, max err %g  (D[0][0..3] = %g %g %g %g)\n", which, bad, 128 * 80, maxerr, D[0],
           D[1], D[2], D[3]);
    if (which != 3 && which != 7) bad_total += bad;      // 3 and 7 are the "other reading" probes
  }
  {
    cudaMemset(dD, 0xff, 128 * 144 * 4);
    probe_kernel<<<1, 128, smem>>>(5, (const __nv_bfloat16*)dU, (const uint8_t*)dW, (const __nv_bfloat16*)dWT, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("case 5: CUDA error %s\n", cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> D(128 * 144);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    double maxerr = 0;
    auto urow = [&](int r, int k) { return U[(r % 80) * NCH + k + 64 * (r / 80)]; };
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 144; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)W[m * 128 + k] * urow(n, k);
        if (n >= 72) ref *= 2;      // the N=80 product from row 72 lands on the same columns
        double err = fabs(ref - D[m * 144 + n]);
        if (!(err <= 1e-3)) ++bad;
        if (err > maxerr || err != err) maxerr = err;
      }
    printf("case 5: mismatches %d / %d, max err %g\n", bad, 128 * 144, maxerr);
    bad_total += bad;
  }
  printf(bad_total == 0 ? "PROBE OK\n" : "PROBE FAILED\n");
  return 0;
}
