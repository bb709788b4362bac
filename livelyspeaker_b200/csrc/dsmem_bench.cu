// Micro-benchmark for the pair design of the fused step kernel (DESIGN.md "pair kernel"): what does it cost to
// write half of every operand tile into the PEER CTA's shared memory?
//   mode 0: local  st.shared.b32            (baseline: 16 warps, one 128-byte wavefront per warp store)
//   mode 1: remote st.shared::cluster.b32   both CTAs of the pair write to each other (bidirectional)
//   mode 2: remote st.shared::cluster.b32   only CTA 0 writes (unidirectional)
//   mode 3: remote st.shared::cluster.v4.b32 (16 bytes per lane), bidirectional
//   mode 4: bulk copy shared::cta -> shared::cluster (peer), 16 KB chunks, completion on the peer's mbarrier,
//           bidirectional
//   mode 5: remote mbarrier arrive ping-pong latency (cycles per round trip / 2)
// Build: make dsmem_bench ; run on the B200 (wrap in `timeout`).
#include <cstdio>
#include <cstdlib>
#include "ls_tc.cuh"
using namespace lstc;

__device__ __forceinline__ void st_cluster_b32(uint32_t caddr, uint32_t v) {
  asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t caddr, uint32_t v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(caddr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_local_b32(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

constexpr int NT = 512;
constexpr uint32_t REGION = 64 * 1024;     // bytes written round-robin

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) dsmem_kernel(int mode, int rounds, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  for (uint32_t i = tid * 4; i < 2 * REGION; i += NT * 4) *reinterpret_cast<uint32_t*>(sm + i) = 0;
  __syncthreads();
  cluster_sync_all();
  const uint32_t base_local = smem_u32(sm);
  const uint32_t base_peer = mapa_u32(base_local, peer);
  long long t0 = clock64();
  if (mode <= 3) {
    const bool active = (mode != 2) || rank == 0;
    const uint32_t base = (mode == 0) ? base_local : base_peer;
    if (active) {
      if (mode == 3) {
        // 16 warps x 512 B per store
        for (int r = 0; r < rounds; ++r)
#pragma unroll 4
          for (uint32_t off = 0; off < REGION; off += 16 * 512)
            st_cluster_v4(base + off + warp * 512 + lane * 16, (uint32_t)r);
      } else if (mode == 0) {
        for (int r = 0; r < rounds; ++r)
#pragma unroll 8
          for (uint32_t off = 0; off < REGION; off += 16 * 128)
            st_local_b32(base + off + warp * 128 + lane * 4, (uint32_t)r);
      } else {
        for (int r = 0; r < rounds; ++r)
#pragma unroll 8
          for (uint32_t off = 0; off < REGION; off += 16 * 128)
            st_cluster_b32(base + off + warp * 128 + lane * 4, (uint32_t)r);
      }
    }
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
    cluster_sync_all();
  } else if (mode == 4) {
    // each CTA: one thread pushes `rounds` x (REGION / 16 KB) chunks into the peer's upper region; the peer's
    // mbarrier counts the bytes
    const uint32_t chunks = REGION / 16384;
    const uint32_t mybar = smem_u32(&bar[0]);
    const uint32_t peerbar = mapa_u32(mybar, peer);
    if (tid == 0) {
      for (int r = 0; r < rounds; ++r) {
        mbar_arrive_expect_tx(&bar[0], REGION);          // bytes that will land HERE from the peer
        for (uint32_t c = 0; c < chunks; ++c)
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           base_peer + REGION + c * 16384),
                       "r"(base_local + c * 16384), "r"(16384u), "r"(peerbar)
                       : "memory");
        mbar_wait(&bar[0], r & 1);
      }
    }
    __syncthreads();
    cluster_sync_all();
  } else {
    // ping-pong: rank 0 arrives on the peer's bar[0], peer waits and arrives on rank 0's bar[1]
    const uint32_t b0 = smem_u32(&bar[0]), b1 = smem_u32(&bar[1]);
    if (tid == 0) {
      for (int r = 0; r < rounds; ++r) {
        if (rank == 0) {
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(b0, 1)) : "memory");
          mbar_wait(&bar[1], r & 1);
        } else {
          mbar_wait(&bar[0], r & 1);
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(b1, 0)) : "memory");
        }
      }
    }
    __syncthreads();
    cluster_sync_all();
  }
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 2 * 148 * 8);
  const int smem = 2 * REGION + 2048;
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"local st.shared.b32 (16 warps)", "remote st.shared::cluster.b32 bidirectional",
                         "remote st.shared::cluster.b32 unidirectional", "remote st.shared::cluster.v4 bidirectional",
                         "bulk copy smem -> peer smem (16 KB chunks) bidirectional", "remote mbarrier arrive ping-pong"};
  for (int grid : {2, 148})
    for (int mode = 0; mode < 6; ++mode) {
      const int rounds = (mode == 5) ? 2000 : 64;
      dsmem_kernel<<<grid, NT, smem>>>(mode, rounds, d);
      dsmem_kernel<<<grid, NT, smem>>>(mode, rounds, d);
      long long c[296];
      cudaError_t e = cudaMemcpy(c, d, grid * 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = c[i] > mx ? c[i] : mx;
      if (mode == 5)
        printf("grid %3d  %-58s: %.0f cycles one way\n", grid, names[mode], (double)mx / rounds / 2);
      else
        printf("grid %3d  %-58s: %.1f B/cycle per writing CTA  (%lld cycles for %d KB)\n", grid, names[mode],
               (double)REGION * rounds / mx, mx, (int)(REGION / 1024) * rounds);
    }
  return 0;
}
