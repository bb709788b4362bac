// Row-tile GEMM on the tcgen05 tensor cores for the once-per-batch projections around the sampling loop (SAG decoder
// layers, the audio half of input_mapping):
//
//   D[row, n] = sum_k A[row, k] * W[n, k]          M = 128 rows per CTA (TMEM lane = row), N = 512 per CTA, fp32 in / out,
//
// bf16x3 split operands (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulation in tensor memory) - the same arithmetic
// as the fused step kernel, so the same parity bar holds.  The A operand is fp32 in global memory: four builder warps
// split it on the fly into the K-major swizzle-128B hi / lo images of a 2-stage ring; W comes
// pre-split from a tape (one 32 KB bulk async copy per [128 n x 64 k] block, 4-slot ring, producer warp); a sixth warp
// issues the MMAs.  Because a thread owns a whole output ROW (its 512 accumulator columns), row-wise epilogues -
// bias + residual + LayerNorm (+ a second LayerNorm after a broadcast add) - need no cross-thread reduction at all.
#pragma once
#include <cuda_bf16.h>

#include "ls_tc.cuh"

namespace lsrg {
using namespace lstc;

constexpr int NT = 512;                          // output columns per CTA = TMEM columns
constexpr int NQ = NT / 128;                     // [128 n x 64 k] weight blocks per K chunk
constexpr uint32_t IMG = 128 * 128;              // 128 rows x 64 bf16, swizzle-128B atoms of 8 rows
constexpr uint32_t STAGE = 2 * IMG;              // hi image, lo image
constexpr int NA = 2, NB = 4;
// Every row tile streams the whole weight matrix, so at 68+ CTAs the L2 -> SM weight traffic (128 KB per CTA per K chunk
// against 3072 cycles of tensor math) is what bounds the kernel (measured: 2x the MMA time).  CL CTAs that work on
// DIFFERENT row tiles of the SAME output columns therefore form a cluster: each fetches 1/CL of every weight stage and
// multicasts it to all of them (the cluster advances in lock step through the weight ring).
constexpr int CL = 1;   // measured: CL = 4 is SLOWER (36 -> 66 us for FFN1 at B = 256: lock step + cluster placement); kept for the record
constexpr uint16_t CL_MASK = (1u << CL) - 1;
constexpr uint32_t OFF_A = 0, OFF_B = NA * STAGE, OFF_BARS = OFF_B + NB * STAGE;
enum { A_FULL0 = 0, A_EMPTY0 = 2, B_FULL0 = 4, B_EMPTY0 = 8, ACC = 12, NBARS = 13 };
constexpr uint32_t OFF_TMEM = OFF_BARS + 16 * 8;
constexpr uint32_t SMEM = OFF_TMEM + 16 + 1024;  // + alignment slack
constexpr int NTHREADS = 192;                    // warps 0-3 build A and run the epilogue, 4 streams W, 5 issues MMAs

__device__ __forceinline__ uint32_t pack_hi_lo(float a, float b, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hb << 16), b - __uint_as_float(hb & 0xFFFF0000u));
  *lo = *reinterpret_cast<const uint32_t*>(&l);
  return hb;
}

// ---- A operand loaders ------------------------------------------------------------------------------------------------
// load(): one 64-wide K chunk of the warp's 32 tile rows into registers; store(): split into bf16 hi / lo and written
// to the swizzled stage.  Both access patterns are COALESCED: a thread-per-row loader (each lane walking its own 2 KB-
// strided row) turns every load into 32 L2 requests and was measured 2x slower once the operands come from DRAM.
struct ARowMajor {               // A[row][k], row stride lda floats (16-byte aligned rows)
  const float* p;
  int lda;
  struct Regs { float4 r[16]; };
  // instruction i: rows 2i, 2i+1 of the warp's 32, lane l -> float4 (l & 15) of row 2i + (l >> 4): 2 x 256 B contiguous
  __device__ __forceinline__ void load(Regs& g, int row0w, int rows, int k0, int lane) const {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int row = row0w + 2 * i + (lane >> 4);
      g.r[i] = row < rows ? __ldg(reinterpret_cast<const float4*>(p + (size_t)row * lda + k0) + (lane & 15))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void store(uint8_t* stage, const Regs& g, int warp, int lane) const {
    const int kq = lane & 15;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int tr = 32 * warp + 2 * i + (lane >> 4);          // tile row
      const uint32_t off = (uint32_t)(tr >> 3) * 1024u + (uint32_t)(tr & 7) * 128u +
                           ((uint32_t)((kq >> 1) ^ (tr & 7)) << 4) + (uint32_t)(kq & 1) * 8u;
      uint2 hi, lo;
      hi.x = pack_hi_lo(g.r[i].x, g.r[i].y, &lo.x);
      hi.y = pack_hi_lo(g.r[i].z, g.r[i].w, &lo.y);
      *reinterpret_cast<uint2*>(stage + off) = hi;              // a half-warp = one row = 128 contiguous (permuted) bytes
      *reinterpret_cast<uint2*>(stage + IMG + off) = lo;
    }
  }
};
struct AChanMajor34 {            // A[clip][k][34 frames]: row = clip * 34 + frame (the WavEncoder's own output layout)
  const float* p;
  int K;
  struct Regs { float v[64]; };
  // lane = row: consecutive rows are consecutive frames, i.e. consecutive addresses for a fixed k
  __device__ __forceinline__ void load(Regs& g, int row0w, int rows, int k0, int lane) const {
    const int row = row0w + lane;
    if (row < rows) {
      const int clip = row / 34, f = row - clip * 34;
      const float* s = p + ((size_t)clip * K + k0) * 34 + f;
#pragma unroll
      for (int i = 0; i < 64; ++i) g.v[i] = __ldg(s + i * 34);
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) g.v[i] = 0.f;
    }
  }
  __device__ __forceinline__ void store(uint8_t* stage, const Regs& g, int warp, int lane) const {
    const int tr = 32 * warp + lane;
    const uint32_t row_off = (uint32_t)(tr >> 3) * 1024u + (uint32_t)(tr & 7) * 128u;
#pragma unroll
    for (int hh = 0; hh < 8; ++hh) {
      uint4 hi, lw;
      hi.x = pack_hi_lo(g.v[8 * hh + 0], g.v[8 * hh + 1], &lw.x);
      hi.y = pack_hi_lo(g.v[8 * hh + 2], g.v[8 * hh + 3], &lw.y);
      hi.z = pack_hi_lo(g.v[8 * hh + 4], g.v[8 * hh + 5], &lw.z);
      hi.w = pack_hi_lo(g.v[8 * hh + 6], g.v[8 * hh + 7], &lw.w);
      const uint32_t off = row_off + ((uint32_t)(hh ^ (tr & 7)) << 4);
      *reinterpret_cast<uint4*>(stage + off) = hi;
      *reinterpret_cast<uint4*>(stage + IMG + off) = lw;
    }
  }
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// ---- epilogues --------------------------------------------------------------------------------------------------------
// Warp w owns tile rows 32w .. 32w+31 (its TMEM lanes); lane = row for everything that touches tensor memory.  Global
// memory is only touched through a per-warp staging tile in shared memory ([32 rows][128 + 4 columns] fp32, in the
// weight ring that is idle by then), 128 columns at a time, so that every global access is a full 512-byte row
// segment per warp instruction - only __syncwarp between the two views, no CTA barrier.
constexpr int STG_LD = 132;                      // floats per staged row (16-byte aligned, conflict-free both ways)
constexpr uint32_t STG_BYTES = 32 * STG_LD * 4;  // per warp

// out[row][n0 + c] = act(acc + bias[n0 + c])
template <bool GELU>
struct EpiStore {
  float* out;
  int ldo;
  const float* bias;             // [N] or nullptr
  __device__ __forceinline__ void run(uint32_t lane_base, int row0w, int rows, int n0, float* stg, int lane) const {
#pragma unroll 1
    for (int cb = 0; cb < NT; cb += 128) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float v[16];
        tmem_ld16(lane_base + cb + 16 * q, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(stg + lane * STG_LD + 16 * q + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      __syncwarp();
      const float4 b4 = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + n0 + cb) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        float4 v = *reinterpret_cast<const float4*>(stg + r * STG_LD + 4 * lane);
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        if (GELU) {
          v.x = gelu_exact(v.x); v.y = gelu_exact(v.y); v.z = gelu_exact(v.z); v.w = gelu_exact(v.w);
        }
        if (row0w + r < rows) *(reinterpret_cast<float4*>(out + (size_t)(row0w + r) * ldo + n0 + cb) + lane) = v;
      }
      __syncwarp();
    }
  }
};

// y = LayerNorm(resid[row] + acc + bias; g1, b1);  TWO: y = LayerNorm(y + add[row / rows_per_add]; g2, b2);  out[row] = y.
// (post-norm nn.TransformerDecoderLayer: norm1 after self-attention, norm2 after the cross-attention whose output is one
// vector per clip here, norm3 after the feed-forward.)  Two-pass statistics like torch's LayerNorm; the row is parked in
// its own accumulator columns between the passes.  out may alias resid (a warp reads its rows before it writes them).
template <bool TWO>
struct EpiResLN {
  float* out;
  const float* resid;            // [rows][512]
  const float* bias;             // [512]
  const float *g1, *b1;
  const float* add;              // [rows / rows_per_add][512]
  int rows_per_add;
  const float *g2, *b2;
  __device__ __forceinline__ float stats(uint32_t lane_base, float sum, float* rstd) const {
    const float mean = sum * (1.f / NT);
    float q = 0.f;
#pragma unroll 2
    for (int c0 = 0; c0 < NT; c0 += 16) {
      float v[16];
      tmem_ld16(lane_base + c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float d = v[j] - mean;
        q = fmaf(d, d, q);
      }
    }
    *rstd = 1.f / sqrtf(q * (1.f / NT) + 1e-5f);
    return mean;
  }
  __device__ __forceinline__ void run(uint32_t lane_base, int row0w, int rows, int, float* stg, int lane) const {
    const int row = row0w + lane;
    const bool valid = row < rows;
    float sum = 0.f;
#pragma unroll 1
    for (int cb = 0; cb < NT; cb += 128) {       // resid + bias, 32 rows x 128 columns through the staging tile
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + cb) + lane);
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0w + r < rows) v = *(reinterpret_cast<const float4*>(resid + (size_t)(row0w + r) * NT + cb) + lane);
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        *reinterpret_cast<float4*>(stg + r * STG_LD + 4 * lane) = v;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float v[16];
        tmem_ld16(lane_base + cb + 16 * q, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 r4 = *reinterpret_cast<const float4*>(stg + lane * STG_LD + 16 * q + j);
          v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) sum += v[j];
        tmem_st16(lane_base + cb + 16 * q, v);
      }
      __syncwarp();
    }
    tmem_st_wait();
    float rstd;
    float mean = stats(lane_base, sum, &rstd);
    if (TWO) {
      const float* ad = add + (size_t)(valid ? row / rows_per_add : 0) * NT;
      sum = 0.f;
#pragma unroll 2
      for (int c0 = 0; c0 < NT; c0 += 16) {
        float v[16];
        tmem_ld16(lane_base + c0, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(g1 + c0 + j)), b = __ldg(reinterpret_cast<const float4*>(b1 + c0 + j)),
                       a = __ldg(reinterpret_cast<const float4*>(ad + c0 + j));
          v[j] = (v[j] - mean) * rstd * g.x + b.x + a.x;
          v[j + 1] = (v[j + 1] - mean) * rstd * g.y + b.y + a.y;
          v[j + 2] = (v[j + 2] - mean) * rstd * g.z + b.z + a.z;
          v[j + 3] = (v[j + 3] - mean) * rstd * g.w + b.w + a.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) sum += v[j];
        tmem_st16(lane_base + c0, v);
      }
      tmem_st_wait();
      mean = stats(lane_base, sum, &rstd);
    }
    const float* g = TWO ? g2 : g1;
    const float* b = TWO ? b2 : b1;
#pragma unroll 1
    for (int cb = 0; cb < NT; cb += 128) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float v[16];
        tmem_ld16(lane_base + cb + 16 * q, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(stg + lane * STG_LD + 16 * q + j) =
              make_float4((v[j] - mean) * rstd, (v[j + 1] - mean) * rstd, (v[j + 2] - mean) * rstd, (v[j + 3] - mean) * rstd);
      }
      __syncwarp();
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + cb) + lane), b4 = __ldg(reinterpret_cast<const float4*>(b + cb) + lane);
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        float4 v = *reinterpret_cast<const float4*>(stg + r * STG_LD + 4 * lane);
        v.x = v.x * g4.x + b4.x; v.y = v.y * g4.y + b4.y; v.z = v.z * g4.z + b4.z; v.w = v.w * g4.w + b4.w;
        if (row0w + r < rows) *(reinterpret_cast<float4*>(out + (size_t)(row0w + r) * NT + cb) + lane) = v;
      }
      __syncwarp();
    }
  }
};

// grid = (row tiles rounded up to a multiple of CL, N / 512); K a multiple of 64.  tape: block (n / 128, k / 64) at ((n / 128) * (K / 64) + k / 64) * STAGE.
template <class ALoad, class Epi>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NTHREADS, 1) rows_gemm_kernel(ALoad al, const uint8_t* __restrict__ tape, int rows, int K,
                                                                Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_chunks = K / 64, n_units = n_chunks * NQ;
  if (tid == 0) {
    for (int i = 0; i < NA; ++i) {
      mbar_init(&bars[A_FULL0 + i], 4);       // one arrival per builder warp
      mbar_init(&bars[A_EMPTY0 + i], 1);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(&bars[B_FULL0 + i], 1);
      mbar_init(&bars[B_EMPTY0 + i], CL);     // the MMA issuers of all CTAs of the cluster release a slot
    }
    mbar_init(&bars[ACC], 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<NT>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();             // barriers of every CTA initialised before any multicast touches them
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sm_s = smem_u32(sm), bars_s = smem_u32(bars);

  if (warp < 4) {
    // ================= A builders, then the epilogue ====================================================
    const int row0w = blockIdx.x * 128 + 32 * warp;            // first row of this warp
    // software pipeline: the global loads of chunk c + 1 are in flight while chunk c is split and stored
    typename ALoad::Regs cur, nxt;
    al.load(nxt, row0w, rows, 0, lane);
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c & 1;
      cur = nxt;
      if (c + 1 < n_chunks) al.load(nxt, row0w, rows, (c + 1) * 64, lane);
      if (c >= NA) {                            // the MMAs that read this stage (chunk c - 2) are done
        mbar_wait(&bars[A_EMPTY0 + s], ((c >> 1) - 1) & 1);
      }
      al.store(sm + OFF_A + s * STAGE, cur, warp, lane);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[A_FULL0 + s]);
    }
    mbar_wait(&bars[ACC], 0);                   // every MMA is complete: the weight ring is free for the staging tiles
    __syncwarp();
    tc_fence_after_sync();
    epi.run(tmem + ((uint32_t)(warp * 32) << 16), row0w, rows, (int)blockIdx.y * NT,
            reinterpret_cast<float*>(sm + OFF_B + warp * STG_BYTES), lane);
    tc_fence_before_sync();
  } else if (warp == 4) {
    // ================= W producer =====================================================================
    // one thread sustains about one bulk copy per 600 cycles whatever its size (bulk_bench.cu): lane j owns ring slot j
    if (lane < NB) {
      const uint32_t rank = cluster_ctarank();
      constexpr uint32_t PART = STAGE / CL;
      const uint8_t* src0 = tape + (size_t)blockIdx.y * NQ * n_chunks * STAGE + rank * PART;
      for (int u = lane; u < n_units; u += NB) {
        const int slot = lane, c = u / NQ, q = u - c * NQ;
        if (u >= NB) mbar_wait_s(bars_s + 8 * (B_EMPTY0 + slot), ((u / NB) - 1) & 1);     // all CTAs are done with the slot
        mbar_arrive_expect_tx_s(bars_s + 8 * (B_FULL0 + slot), STAGE);                    // CL parts, one from each CTA
        bulk_g2s_mc_s(sm_s + OFF_B + slot * STAGE + rank * PART, src0 + ((size_t)q * n_chunks + c) * STAGE, PART,
                      bars_s + 8 * (B_FULL0 + slot), CL_MASK);
      }
    }
  } else {
    // ================= MMA issuer: the whole warp in uniform control flow, one elected lane issues ======
    constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
    constexpr uint32_t idesc = idesc_bf16(128, 128, 0, 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    int u = 0;
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c & 1;
      mbar_wait_s(bars_s + 8 * (A_FULL0 + s), (c >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t a_hi = desc_lo32(sm_s + OFF_A + s * STAGE, 16), a_lo = desc_lo32(sm_s + OFF_A + s * STAGE + IMG, 16);
#pragma unroll 1
      for (int q = 0; q < NQ; ++q, ++u) {
        const int slot = u % NB;
        mbar_wait_s(bars_s + 8 * (B_FULL0 + slot), (u / NB) & 1);
        tc_fence_after_sync();
        const uint32_t b_hi = desc_lo32(sm_s + OFF_B + slot * STAGE, 16), b_lo = desc_lo32(sm_s + OFF_B + slot * STAGE + IMG, 16);
        const uint32_t d = tm + (uint32_t)q * 128u;
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          umma_bf16_split_elect(d, a_hi + 2 * ks, DH, b_hi + 2 * ks, DH, idesc, (c > 0 || ks > 0) ? 1u : 0u);
          umma_bf16_split_elect(d, a_lo + 2 * ks, DH, b_hi + 2 * ks, DH, idesc, 1u);
          umma_bf16_split_elect(d, a_hi + 2 * ks, DH, b_lo + 2 * ks, DH, idesc, 1u);
        }
        umma_commit_mc_s_elect(bars_s + 8 * (B_EMPTY0 + slot), CL_MASK);
      }
      umma_commit_s_elect(bars_s + 8 * (A_EMPTY0 + s));
    }
    umma_commit_s_elect(bars_s + 8 * ACC);
  }
  cluster_sync_all();             // peers may still multicast commits into this CTA's barriers
  if (warp == 5) tmem_dealloc<NT>(tmem);
}

// tape of W given TRANSPOSED (wt[k * N + n], fp32): zero padding beyond N / K
static __global__ void build_rows_tape_kernel(const float* __restrict__ wt, int N, int K, uint8_t* __restrict__ dst) {
  const int nblk = (N + 127) / 128, kblk = K / 64;
  const long long total = (long long)nblk * kblk * 128 * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int nl = (int)(i & 127), kl = (int)((i >> 7) & 63);
    const long long blk = i >> 13;
    const int kb = (int)(blk % kblk), nb = (int)(blk / kblk);
    const int n = nb * 128 + nl, k = kb * 64 + kl;
    const float v = n < N ? wt[(size_t)k * N + n] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    uint8_t* base = dst + (size_t)blk * STAGE;
    const uint32_t off = tile_off(nl, kl, 0);
    *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + IMG + off) = lo;
  }
}
inline int row_tiles(int rows) { return ((rows + 127) / 128 + CL - 1) / CL * CL; }
inline size_t rows_tape_bytes(int N, int K) { return (size_t)((N + 127) / 128) * (K / 64) * STAGE; }

}  // namespace lsrg
