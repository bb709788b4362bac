// Micro-benchmark 4 (ring depth variant of benchmark 3; CG = 1 only): the GEMM phase of the fused kernel with its weight ring, single CTA vs CTA pair.
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



template <int CG, int FILL, int NSLOT, uint32_t SLOT, int SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1)
bench3_kernel(const uint8_t* __restrict__ src, int pairs, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  // [0, 64K) ring, [64K, 64K + 147456) operand tile U: 8 blocks x (lo image 9 KB | hi image 9 KB)
  __shared__ uint64_t bars[2 * NSLOT + 1];     // full[NSLOT], empty[NSLOT], done
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  for (uint32_t i = tid * 4; i < NSLOT * SLOT + 2 * 18432 + 2048; i += 160 * 4) *reinterpret_cast<uint32_t*>(sm + i) = 0x3c003c00u;
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&bars[NSLOT + s], 1);
    }
    mbar_init(&bars[2 * NSLOT], 1);
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
  const uint32_t bars_s = smem_u32(bars), ring_s = smem_u32(sm), u_s = smem_u32(sm + NSLOT * SLOT);
  constexpr uint32_t KPS = SLOT / 4096;   // K steps (MMAs) per slot
  const uint32_t stages = 2u * (uint32_t)pairs * (4 / KPS);
  const long long t0 = clock64();
  if (warp == 1) {
    // ---- producer: lane j owns slot j ----
    // SPLIT lanes per slot: lane = slot * SPLIT + part; every part copies SLOT / SPLIT bytes, part 0 posts the expected bytes
    if (lane < NSLOT * SPLIT) {
      const uint32_t slot = lane / SPLIT, part = lane % SPLIT;
      constexpr uint32_t PB = SLOT / SPLIT;
      for (uint32_t it = slot; it < stages; it += NSLOT) {
        mbar_wait_s(bars_s + 8 * (NSLOT + slot), ((it / NSLOT) & 1) ^ 1);
        if (FILL) {
          if (part == 0) mbar_arrive_expect_tx_s(bars_s + 8 * slot, SLOT);
          bulk_g2s_s(ring_s + slot * SLOT + part * PB, src + ((size_t)(blockIdx.x * 131 + it) % 512) * 16384 + part * PB, PB,
                     bars_s + 8 * slot);
        } else if (part == 0) {
          mbar_arrive_expect_tx_s(bars_s + 8 * slot, 0);
        }
      }
    }
  } else if (warp == 2 && CG == 2 && rank == 1) {
    // ---- relay in the peer: stage landed here -> tell the leader ----
    if (lane < NSLOT) {
      const uint32_t remote = mapa_u32(bars_s + 8 * lane, 0);
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
      
      tc_fence_after_sync();
      return desc_lo32(ring_s + slot * SLOT, 16);
    };
    auto release = [&]() {
      umma_commit_s_elect(bars_s + 8 * (NSLOT + (it % NSLOT)));
      ++it;
    };
#pragma unroll 1
    for (int p = 0; p < pairs; ++p) {
      const uint32_t ub = uk + (uint32_t)(p & 1) * BLK;
      const uint32_t d = tmem + (uint32_t)((p >> 3) & 1) * 152;
#pragma unroll
      for (uint32_t sub = 0; sub < 4 / KPS; ++sub) {     // hi block: N = 144
        const uint32_t wl = wait_stage();
#pragma unroll
        for (uint32_t ks = 0; ks < KPS; ++ks)
          umma_bf16_split_elect(d, wl + 2 * ks, DH, ub + 2 * (sub * KPS + ks), DH, id144, 1u);
        release();
      }
#pragma unroll
      for (uint32_t sub = 0; sub < 4 / KPS; ++sub) {     // lo block: N = 80
        const uint32_t wl = wait_stage();
#pragma unroll
        for (uint32_t ks = 0; ks < KPS; ++ks)
          umma_bf16_split_elect(d + 72, wl + 2 * ks, DH, ub + HI + 2 * (sub * KPS + ks), DH, id80, 1u);
        release();
      }
    }
    umma_commit_s_elect(bars_s + 8 * (2 * NSLOT));
  }
  if (warp == 4) {
    mbar_wait_s(bars_s + 8 * (2 * NSLOT), 0);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) {
    if (CG == 2) tmem_dealloc2<512>(tmem);
    else tmem_dealloc<512>(tmem);
  }
}

template <int CG, int FILL, int NSLOT, uint32_t SLOT, int SPLIT = 1>
void run(const uint8_t* src, long long* d, int grid) {
  const int smem = NSLOT * SLOT + 2 * 18432 + 2048 + 1024, pairs = 2048;
  cudaFuncSetAttribute(bench3_kernel<CG, FILL, NSLOT, SLOT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench3_kernel<CG, FILL, NSLOT, SLOT, SPLIT><<<grid, 160, smem>>>(src, pairs, d);
  bench3_kernel<CG, FILL, NSLOT, SLOT, SPLIT><<<grid, 160, smem>>>(src, pairs, d);
  long long c[148];
  cudaError_t e = cudaMemcpy(c, d, grid * 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error CG=%d FILL=%d: %s\n", CG, FILL, cudaGetErrorString(e)); exit(1); }
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = c[i] > mx ? c[i] : mx;
  const double per_pair = (double)mx / pairs;
  // clip-equivalents per CTA per weight-block pair: CG=1 one clip x 128 channels; CG=2 two clips x 128 channels
  printf("grid %3d  ring %2d x %2d KB (%d copies per slot)  fill %d : %.0f cycles per weight-block pair (tensor math alone: 496)\n", grid, NSLOT, (int)(SLOT / 1024), SPLIT, FILL, per_pair);
}

int main() {
  uint8_t* src;
  long long* d;
  cudaMalloc(&src, 512 * 16384);
  cudaMemset(src, 0x3c, 512 * 16384);
  cudaMalloc(&d, 148 * 8);
  for (int grid : {2, 148}) {
    run<1, 0, 4, 16384>(src, d, grid);
    run<1, 1, 2, 16384>(src, d, grid);
    run<1, 1, 3, 16384>(src, d, grid);
    run<1, 1, 4, 16384>(src, d, grid);
    run<1, 1, 6, 16384>(src, d, grid);
    run<1, 1, 8, 16384>(src, d, grid);
    run<1, 1, 10, 16384>(src, d, grid);
    run<1, 1, 4, 16384, 2>(src, d, grid);
    run<1, 1, 4, 16384, 4>(src, d, grid);
    run<1, 1, 4, 16384, 8>(src, d, grid);
    run<1, 1, 5, 16384, 4>(src, d, grid);
    run<1, 1, 6, 16384, 4>(src, d, grid);
    run<1, 0, 8, 8192>(src, d, grid);
    run<1, 1, 8, 8192>(src, d, grid);
    run<1, 1, 9, 8192>(src, d, grid);
    run<1, 0, 16, 4096>(src, d, grid);
    run<1, 1, 16, 4096>(src, d, grid);
    run<1, 1, 12, 4096>(src, d, grid);
    run<1, 1, 18, 4096>(src, d, grid);
  }
  return 0;
}
