// Micro-benchmark 3: the GEMM phase of the fused kernel with its weight ring, single CTA vs CTA pair.
//   CG = 1: every CTA on its own (today's kernel): per weight-block pair 4 x (M=128, N=144) MMAs on the hi stage and
//           4 x (M=128, N=80) on the lo stage, each 16 KB stage streamed from L2 into a 4-slot ring by bulk copies.
//   CG = 2: cluster pair, cta_group::2 (M=256, N=144 = 72 rows of B from each CTA): 8 MMAs on the hi stage (x U_hi,
//           x U_lo), 4 on the lo stage; each CTA streams ITS 16 KB stages; the leader issues, a relay warp in the peer
//           forwards "stage landed" to the leader, commits free the slot in both CTAs (multicast).
//   FILL = 0: stages are not copied (pure MMA + operand reads).
// Prints cycles per weight-block pair (hi + lo stage), per CTA.  Build: make umma_bench3 ; run under `timeout`.
#include <cstdio>
#include <cstdlib>
#include "ls_tc.cuh"
using namespace lstc;

constexpr uint32_t SLOT = 16384;
constexpr int NSLOT = 4;

template <int CG, int FILL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1)
bench3_kernel(const uint8_t* __restrict__ src, int pairs, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  // [0, 64K) ring, [64K, 64K + 147456) operand tile U: 8 blocks x (lo image 9 KB | hi image 9 KB)
  __shared__ uint64_t bars[16];     // full[4], empty[4], pfull[4], done
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  for (uint32_t i = tid * 4; i < 64 * 1024 + 147456; i += 160 * 4) *reinterpret_cast<uint32_t*>(sm + i) = 0x3c003c00u;
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&bars[4 + s], 1);
      mbar_init(&bars[8 + s], 1);
    }
    mbar_init(&bars[12], 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    if (CG == 2) tmem_alloc2<512>(&tslot);
    else tmem_alloc<512>(&tslot);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tslot, 0);
  const uint32_t bars_s = smem_u32(bars), ring_s = smem_u32(sm), u_s = smem_u32(sm + 64 * 1024);
  const uint32_t stages = 2u * (uint32_t)pairs;
  const long long t0 = clock64();
  if (warp == 1) {
    // ---- producer: lane j owns slot j ----
    if (lane < NSLOT) {
      for (uint32_t it = lane; it < stages; it += NSLOT) {
        mbar_wait_s(bars_s + 8 * (4 + lane), ((it / NSLOT) & 1) ^ 1);
        if (FILL) {
          mbar_arrive_expect_tx_s(bars_s + 8 * lane, SLOT);
          bulk_g2s_s(ring_s + lane * SLOT, src + ((size_t)(blockIdx.x * 131 + it) % 512) * SLOT, SLOT, bars_s + 8 * lane);
        } else {
          mbar_arrive_expect_tx_s(bars_s + 8 * lane, 0);
        }
      }
    }
  } else if (warp == 2 && CG == 2 && rank == 1) {
    // ---- relay in the peer: stage landed here -> tell the leader ----
    if (lane < NSLOT) {
      const uint32_t remote = mapa_u32(bars_s + 8 * (8 + lane), 0);
      for (uint32_t it = lane; it < stages; it += NSLOT) {
        mbar_wait_s(bars_s + 8 * lane, (it / NSLOT) & 1);
        mbar_arrive_cluster(remote);
      }
    }
  } else if (warp == 3 && (CG == 1 || rank == 0)) {
    // ---- MMA issuer ----
    constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
    constexpr uint32_t id144 = idesc_bf16(CG == 2 ? 256 : 128, 144, 0, 0), id80 = idesc_bf16(128, 80, 0, 0);
    const uint32_t uk = desc_lo32(u_s, 16);
    constexpr uint32_t HI = 9216 >> 4, BLK = 18432 >> 4;
    uint32_t it = 0;
    auto wait_stage = [&]() -> uint32_t {
      const uint32_t slot = it % NSLOT, ph = (it / NSLOT) & 1;
      mbar_wait_s(bars_s + 8 * slot, ph);
      if (CG == 2) mbar_wait_s(bars_s + 8 * (8 + slot), ph);
      tc_fence_after_sync();
      return desc_lo32(ring_s + slot * SLOT, 16);
    };
    auto release = [&]() {
      if (CG == 2) umma2_commit_mc_s_elect(bars_s + 8 * (4 + (it % NSLOT)), (uint16_t)3);
      else umma_commit_s_elect(bars_s + 8 * (4 + (it % NSLOT)));
      ++it;
    };
#pragma unroll 1
    for (int p = 0; p < pairs; ++p) {
      const uint32_t ub = uk + (uint32_t)(p & 7) * BLK;
      const uint32_t d = tmem + (uint32_t)((p >> 3) & 1) * 152;
      uint32_t wl = wait_stage();
      if (CG == 2) {
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) umma2_bf16_split_elect(d, wl + 2 * ks, DH, ub + HI + 2 * ks, DH, id144, 1u);
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) umma2_bf16_split_elect(d, wl + 2 * ks, DH, ub + 2 * ks, DH, id144, 1u);
        release();
        wl = wait_stage();
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) umma2_bf16_split_elect(d, wl + 2 * ks, DH, ub + HI + 2 * ks, DH, id144, 1u);
        release();
      } else {
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) umma_bf16_split_elect(d, wl + 2 * ks, DH, ub + 2 * ks, DH, id144, 1u);
        release();
        wl = wait_stage();
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) umma_bf16_split_elect(d + 72, wl + 2 * ks, DH, ub + HI + 2 * ks, DH, id80, 1u);
        release();
      }
    }
    if (CG == 2) umma2_commit_mc_s_elect(bars_s + 8 * 12, (uint16_t)3);
    else umma_commit_s_elect(bars_s + 8 * 12);
  }
  if (warp == 4) {
    mbar_wait_s(bars_s + 8 * 12, 0);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) {
    if (CG == 2) tmem_dealloc2<512>(tmem);
    else tmem_dealloc<512>(tmem);
  }
}

template <int CG, int FILL>
void run(const uint8_t* src, long long* d, int grid) {
  const int smem = 64 * 1024 + 147456 + 1024 + 2048, pairs = 2048;   // + the N=80 overhang past the last hi image
  cudaFuncSetAttribute(bench3_kernel<CG, FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench3_kernel<CG, FILL><<<grid, 160, smem>>>(src, pairs, d);
  bench3_kernel<CG, FILL><<<grid, 160, smem>>>(src, pairs, d);
  long long c[148];
  cudaError_t e = cudaMemcpy(c, d, grid * 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error CG=%d FILL=%d: %s\n", CG, FILL, cudaGetErrorString(e)); exit(1); }
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = c[i] > mx ? c[i] : mx;
  const double per_pair = (double)mx / pairs;
  // clip-equivalents per CTA per weight-block pair: CG=1 one clip x 128 channels; CG=2 two clips x 128 channels
  printf("grid %3d  cta_group::%d  fill %d : %.0f cycles per weight-block pair per CTA = %.0f per clip  (tensor math alone: %d)\n",
         grid, CG, FILL, per_pair, per_pair / CG, CG == 2 ? 864 : 496);
}

int main() {
  uint8_t* src;
  long long* d;
  cudaMalloc(&src, 512 * SLOT);
  cudaMemset(src, 0x3c, 512 * SLOT);
  cudaMalloc(&d, 148 * 8);
  for (int grid : {2, 148}) {
    run<1, 0>(src, d, grid);
    run<1, 1>(src, d, grid);
    run<2, 0>(src, d, grid);
    run<2, 1>(src, d, grid);
  }
  return 0;
}
