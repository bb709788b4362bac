// Row-tile GEMM on the tcgen05 tensor cores for the once-per-batch projections around the sampling loop (SAG decoder
// layers, the audio half of input_mapping):
//
//   D[row, n] = sum_k A[row, k] * W[n, k]          M = 128 rows per CTA (TMEM lane = row), N = 512 per CTA, fp32 in / out,
//
// bf16x3 split operands (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulation in tensor memory) - the same arithmetic
// as the fused step kernel, so the same parity bar holds.  The A operand is fp32 in global memory: four builder warps
// (thread = row) split it on the fly into the K-major swizzle-128B hi / lo images of a 2-stage ring; W comes
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
constexpr uint32_t OFF_A = 0, OFF_B = NA * STAGE, OFF_BARS = OFF_B + NB * STAGE;
enum { A_FULL0 = 0, A_EMPTY0 = 2, B_FULL0 = 4, B_EMPTY0 = 8, ACC = 12, NBARS = 13 };
constexpr uint32_t OFF_TMEM = OFF_BARS + 16 * 8;
constexpr uint32_t SMEM = OFF_TMEM + 16 + 1024;  // + alignment slack
constexpr int NTHREADS = 192;                    // warps 0-3 build A and run the epilogue, 4 streams W, 5 issues MMAs

// ---- A operand loaders: 64 consecutive k of one row ------------------------------------------------------------------
struct ARowMajor {               // A[row][k], row stride lda floats (16-byte aligned rows)
  const float* p;
  int lda;
  __device__ __forceinline__ void load64(int row, int k0, float (&v)[64]) const {
    const float4* s = reinterpret_cast<const float4*>(p + (size_t)row * lda + k0);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 t = __ldg(s + i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  }
};
struct AChanMajor34 {            // A[clip][k][34 frames]: row = clip * 34 + frame (the WavEncoder's own output layout)
  const float* p;
  int K;
  __device__ __forceinline__ void load64(int row, int k0, float (&v)[64]) const {
    const int clip = row / 34, f = row - clip * 34;
    const float* s = p + ((size_t)clip * K + k0) * 34 + f;
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __ldg(s + i * 34);
  }
};

__device__ __forceinline__ uint32_t pack_hi_lo(float a, float b, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hb << 16), b - __uint_as_float(hb & 0xFFFF0000u));
  *lo = *reinterpret_cast<const uint32_t*>(&l);
  return hb;
}

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

// ---- epilogues: thread = row, `lane_base` = TMEM address of its lane, columns [0, 512) = this CTA's outputs ----------
// out[row][n0 + c] = act(acc + bias[n0 + c])
template <bool GELU>
struct EpiStore {
  float* out;
  int ldo;
  const float* bias;             // [N] or nullptr
  __device__ __forceinline__ void run(uint32_t lane_base, int row, bool valid, int n0) const {
    float* dst = out + (size_t)row * ldo + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 16) {
      float v[16];
      tmem_ld16(lane_base + c0, v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (bias != nullptr) v[j] += __ldg(bias + n0 + c0 + j);
          if (GELU) v[j] = gelu_exact(v[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
};

// y = LayerNorm(resid[row] + acc + bias; g1, b1);  TWO: y = LayerNorm(y + add[row / rows_per_add]; g2, b2);  out[row] = y.
// (post-norm nn.TransformerDecoderLayer: norm1 after self-attention, norm2 after the cross-attention whose output is one
// vector per clip here, norm3 after the feed-forward.)  Two-pass statistics like torch's LayerNorm; the row is parked in
// its own accumulator columns between the passes.  out may alias resid (each thread reads its row before writing it).
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
#pragma unroll 1
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
  __device__ __forceinline__ void run(uint32_t lane_base, int row, bool valid, int) const {
    const float* rs = resid + (size_t)row * NT;
    float sum = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 16) {
      float v[16];
      tmem_ld16(lane_base + c0, v);
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 r = valid ? *reinterpret_cast<const float4*>(rs + c0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[j] += r.x + __ldg(bias + c0 + j);
        v[j + 1] += r.y + __ldg(bias + c0 + j + 1);
        v[j + 2] += r.z + __ldg(bias + c0 + j + 2);
        v[j + 3] += r.w + __ldg(bias + c0 + j + 3);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) sum += v[j];
      tmem_st16(lane_base + c0, v);
    }
    tmem_st_wait();
    float rstd;
    float mean = stats(lane_base, sum, &rstd);
    if (TWO) {
      const float* ad = add + (size_t)(valid ? row / rows_per_add : 0) * NT;
      sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 16) {
        float v[16];
        tmem_ld16(lane_base + c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = (v[j] - mean) * rstd * __ldg(g1 + c0 + j) + __ldg(b1 + c0 + j) + __ldg(ad + c0 + j);
          sum += v[j];
        }
        tmem_st16(lane_base + c0, v);
      }
      tmem_st_wait();
      mean = stats(lane_base, sum, &rstd);
    }
    const float* g = TWO ? g2 : g1;
    const float* b = TWO ? b2 : b1;
    float* dst = out + (size_t)row * NT;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 16) {
      float v[16];
      tmem_ld16(lane_base + c0, v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (v[j] - mean) * rstd * __ldg(g + c0 + j) + __ldg(b + c0 + j);
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
};

// grid = (row tiles, N / 512); K a multiple of 64.  tape: block (n / 128, k / 64) at ((n / 128) * (K / 64) + k / 64) * STAGE.
template <class ALoad, class Epi>
__global__ void __launch_bounds__(NTHREADS, 1) rows_gemm_kernel(ALoad al, const uint8_t* __restrict__ tape, int rows, int K,
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
      mbar_init(&bars[B_EMPTY0 + i], 1);
    }
    mbar_init(&bars[ACC], 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<NT>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sm_s = smem_u32(sm), bars_s = smem_u32(bars);

  if (warp < 4) {
    // ================= A builders, then the epilogue: thread = row =====================================
    const int row = blockIdx.x * 128 + tid;
    const bool valid = row < rows;
    const uint32_t row_off = (uint32_t)(tid >> 3) * 1024u + (uint32_t)(tid & 7) * 128u;
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c & 1;
      float v[64];
      if (valid) {
        al.load64(row, c * 64, v);
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = 0.f;
      }
      if (c >= NA) {                            // the MMAs that read this stage (chunk c - 2) are done
        mbar_wait(&bars[A_EMPTY0 + s], ((c >> 1) - 1) & 1);
      }
      uint8_t* stage = sm + OFF_A + s * STAGE;
#pragma unroll
      for (int hh = 0; hh < 8; ++hh) {
        uint4 hi, lw;
        hi.x = pack_hi_lo(v[8 * hh + 0], v[8 * hh + 1], &lw.x);
        hi.y = pack_hi_lo(v[8 * hh + 2], v[8 * hh + 3], &lw.y);
        hi.z = pack_hi_lo(v[8 * hh + 4], v[8 * hh + 5], &lw.z);
        hi.w = pack_hi_lo(v[8 * hh + 6], v[8 * hh + 7], &lw.w);
        const uint32_t off = row_off + ((uint32_t)(hh ^ (tid & 7)) << 4);
        *reinterpret_cast<uint4*>(stage + off) = hi;
        *reinterpret_cast<uint4*>(stage + IMG + off) = lw;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[A_FULL0 + s]);
    }
    mbar_wait(&bars[ACC], 0);
    __syncwarp();
    tc_fence_after_sync();
    epi.run(tmem + ((uint32_t)(warp * 32) << 16), row, valid, (int)blockIdx.y * NT);
    tc_fence_before_sync();
  } else if (warp == 4) {
    // ================= W producer =====================================================================
    if (lane == 0) {
      const uint8_t* src0 = tape + (size_t)blockIdx.y * NQ * n_chunks * STAGE;
      for (int u = 0; u < n_units; ++u) {
        const int slot = u % NB, c = u / NQ, q = u - c * NQ;
        if (u >= NB) mbar_wait_s(bars_s + 8 * (B_EMPTY0 + slot), ((u / NB) - 1) & 1);
        mbar_arrive_expect_tx_s(bars_s + 8 * (B_FULL0 + slot), STAGE);
        bulk_g2s_s(sm_s + OFF_B + slot * STAGE, src0 + ((size_t)q * n_chunks + c) * STAGE, STAGE, bars_s + 8 * (B_FULL0 + slot));
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
        umma_commit_s_elect(bars_s + 8 * (B_EMPTY0 + slot));
      }
      umma_commit_s_elect(bars_s + 8 * (A_EMPTY0 + s));
    }
    umma_commit_s_elect(bars_s + 8 * ACC);
  }
  __syncthreads();
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
inline size_t rows_tape_bytes(int N, int K) { return (size_t)((N + 127) / 128) * (K / 64) * STAGE; }

}  // namespace lsrg
