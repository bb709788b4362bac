// The fused denoising step: ONE persistent sm_100a kernel per step.
//
//   _WrappedModel + ClassifierFreeSampleModel + RAG.forward (minus the hoisted terms) +
//   p_mean_variance + p_sample / ddim_sample of the reference
//   (respace.py:118-130, cfg_sampler.py:24-31, RAG.py:98-133, mlp_module.py:37-91,
//    gaussian_diffusion.py:284-399, 507-558, 745-798) - see DESIGN.md "fused step kernel".
//
// Work unit (tile) = one clip, both guidance passes: 2*S token rows (S = 35 TED / 36 BEAT).
// All GEMMs run TRANSPOSED on the tensor cores, D^T[channel, row] = W[channel, k] * U^T[k, row]:
//   M = 128 output channels (4 M-tiles for d = 512), N = 80 rows (2*S padded), K = channels.
// so the TMEM lane of an accumulator element is its CHANNEL and the column is its ROW:
//   * each of the 512 epilogue threads owns one channel and keeps the fp32 residual stream
//     h[row] of that channel in REGISTERS for the whole 8-layer stack;
//   * TMEM holds only accumulators (4 M-tiles x 80 columns);
//   * shared memory holds the bf16 hi/lo operand tile U (LayerNorm output) - the very same
//     bytes serve as the K-major B operand of the channel-mix GEMM and as the MN-major A
//     operand of the token-mix GEMM (ls_tc.cuh, validated by umma_probe.cu) - plus a
//     2-slot ring of weight stages streamed from L2 by 1-D bulk async copies.
// Precision: PRECISE = bf16x3 (hi*hi + lo*hi + hi*lo, fp32 accumulate) meets the
// rtol 1e-3 / atol 1e-4 parity bar; !PRECISE = plain bf16 operands (fast mode).
//
// Warp roles (20 warps): 0-15 epilogue (thread id == channel), 16 weight producer,
// 17 MMA issuer + TMEM owner, 18-19 idle; warps 16-19 give most of their registers to the
// epilogue warps with setmaxnreg so the 70-row residual stream fits without spilling.
#include <cuda_bf16.h>
#include <cstdlib>

#include "ls_internal.cuh"
#include "ls_tc.cuh"
#include "ls_update.cuh"

using namespace lstc;

namespace {

constexpr int NT_EPI = 512;
constexpr int NT_ALL = 640;                     // 16 epilogue warps + 1 producer + 1 MMA + 2 idle (register donors)
constexpr int REGS_EPI = 112;                   // setmaxnreg budgets must conserve the launch allocation:
constexpr int REGS_AUX = 32;                    // 512*112 + 128*32 == 640*96 (a larger total blocks forever in setmaxnreg.inc)
constexpr int NROW = 80;                       // MMA N (rows of a tile, padded)
constexpr int RGS = 9;                         // 8-row groups stored per 64-channel block (72 rows)
constexpr uint32_t CBS = RGS * 1024;           // bytes between 64-channel blocks of U
constexpr uint32_t U_BYTES = 8 * CBS;          // one of U_hi / U_lo
constexpr uint32_t SLOT = 16384;               // ring slot = tape stage pitch
constexpr int NSLOT = 4;
#ifndef LS_MULTICAST
#define LS_MULTICAST 1   // 1: weight stages fetched half by each CTA of the pair and multicast to both
#endif
constexpr uint32_t W_HALF = 16384;             // 128 x 64 bf16 weight block (hi or lo) = one stage
constexpr uint32_t WBLK_BLK = 10 * 1024;       // token-mix weights: 80 rows x 64 k = one stage

// shared memory map (offsets from a 1024-aligned base)
constexpr uint32_t OFF_UHI = 0;
constexpr uint32_t OFF_ULO = U_BYTES;
constexpr uint32_t OFF_PAD = 2 * U_BYTES;               // 1024 zero bytes: row group 9 of the last block
constexpr uint32_t OFF_RING = OFF_PAD + 1024;
constexpr uint32_t OFF_PART = OFF_RING + NSLOT * SLOT;  // [16 warps][72 rows] float2 LN partial sums
constexpr uint32_t OFF_STATS = OFF_PART + 16 * 72 * 8;  // [72] float2 (mean, rstd)
constexpr uint32_t OFF_BTOK = OFF_STATS + 72 * 8;       // [72] float token-mix bias
constexpr uint32_t OFF_BARS = OFF_BTOK + 72 * 4;        // mbarriers
constexpr uint32_t OFF_TMEM = OFF_BARS + 16 * 8;
constexpr uint32_t SMEM_USED = OFF_TMEM + 16;
constexpr uint32_t SMEM_DYN = SMEM_USED + 1024;         // + alignment slack

enum { BAR_FULL0 = 0, BAR_EMPTY0 = 4, BAR_UREADY0 = 8, BAR_ACC0 = 12 };   // indices into the mbarrier array

struct FusedParams {
  const uint8_t* tape;     // weight stages, SLOT bytes apart
  int n_layers, JD, KIN, MH, B;
  LsWeights w;
  const float* A; const float* P; const float* z_mu; const float* z_lv; const float* emo_tok;
  const float* x_t; const float* eps_c; const float* eps_u; const float* noise; const float* scale;
  long long nsb, nsj, nsf;
  float* x_prev; float* pred_x0;
  ls_step_params sp;
  long long* timing;       // debug (LS_FUSED_TIMING=1): clock64 stamps of block 0, threads 0 and 511
};

// x * sigmoid(x) with ex2.approx + rcp.approx (about 2 ulp; the IEEE reciprocal costs ~10 more instructions)
__device__ __forceinline__ float silu_fast(float z) { return __fdividef(z, 1.f + __expf(-z)); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// hi/lo bf16 split of v, stored at byte offset `off` of U_hi (and U_lo)
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(v) : "memory");
}
// u_s = shared-space address of U_hi
template <bool PRECISE>
__device__ __forceinline__ void store_split(uint32_t u_s, uint32_t off, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  sts_u16(u_s + off, __bfloat16_as_ushort(hi));
  if (PRECISE) {
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    sts_u16(u_s + OFF_ULO + off, __bfloat16_as_ushort(lo));
  }
}

// Two values (rows n0, n1 of this thread's channel) -> bf16 hi (+ lo) with the packed converter
// (F2FP, full-rate pipe; the scalar F2F conversion issues at MUFU rate and was the LayerNorm
// store bottleneck).
template <bool PRECISE>
__device__ __forceinline__ void store_split2(uint32_t u_s, uint32_t off0, uint32_t off1, float v0, float v1) {
  const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);      // .x = v0 (low half), .y = v1 (high half)
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hi);
  sts_u16(u_s + off0, (uint16_t)(hb & 0xFFFFu));
  sts_u16(u_s + off1, (uint16_t)(hb >> 16));
  if (PRECISE) {
    const float r0 = v0 - __uint_as_float(hb << 16), r1 = v1 - __uint_as_float(hb & 0xFFFF0000u);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(r0, r1);
    const uint32_t lb = *reinterpret_cast<const uint32_t*>(&lo);
    sts_u16(u_s + OFF_ULO + off0, (uint16_t)(lb & 0xFFFFu));
    sts_u16(u_s + OFF_ULO + off1, (uint16_t)(lb >> 16));
  }
}

__device__ __forceinline__ uint32_t row_off(uint32_t pre_off, int n) {
  return (pre_off ^ ((uint32_t)(n & 7) << 4)) + (uint32_t)(n & 7) * 128u + (uint32_t)(n >> 3) * 1024u;
}

// Per-row (sum, sum of squares) over the 512 channels -> stats[row] = (mean, 1/std).
// h[] holds this thread's channel; `shift` rows come from the previous stats (robust
// single-pass variance), or 0 when use_shift is false.
template <int R>
__device__ __forceinline__ void ln_stats(const float (&h)[72], uint8_t* sm, bool use_shift) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float2* part = reinterpret_cast<float2*>(sm + OFF_PART);
  float2* stats = reinterpret_cast<float2*>(sm + OFF_STATS);
#pragma unroll
  for (int g = 0; g < 9; ++g) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = 8 * g + j;
      float d = 0.f;
      if (n < R) d = h[n] - (use_shift ? stats[n].x : 0.f);
      s[j] = d;
      q[j] = d * d;
    }
    // halving butterfly: after the three steps lane bits (4,3,2) select the row
#pragma unroll
    for (int step = 0; step < 3; ++step) {
      const int half = 4 >> step;            // 4, 2, 1 values kept
      const int xm = 16 >> step;             // xor 16, 8, 4
      const bool up = (lane & xm) != 0;
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float send_s = up ? s[j] : s[j + half];
        const float send_q = up ? q[j] : q[j + half];
        const float keep_s = up ? s[j + half] : s[j];
        const float keep_q = up ? q[j + half] : q[j];
        s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, xm);
        q[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, xm);
      }
    }
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 2);
    q[0] += __shfl_xor_sync(0xffffffffu, q[0], 2);
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
    q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
    if ((lane & 3) == 0) {
      const int row = 8 * g + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      part[warp * 72 + row] = make_float2(s[0], q[0]);
    }
  }
  epi_bar();
  if (tid < R) {              // 70 threads in parallel, 16 pipelined loads each (a warp-shuffle finish was slower)
    float ss = 0.f, qq = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      const float2 p = part[w * 72 + tid];
      ss += p.x;
      qq += p.y;
    }
    const float sh = use_shift ? stats[tid].x : 0.f;
    const float md = ss * (1.f / 512.f);
    const float var = fmaxf(qq * (1.f / 512.f) - md * md, 0.f);
    stats[tid] = make_float2(sh + md, rsqrtf(var + 1e-5f));
  }
  epi_bar();
}

// LayerNorm of h -> bf16 (hi, lo) operand tile.  pre = smem address of (row 0, channel c)
// with the chunk bits at [4,7): row n lives at (pre ^ ((n&7)<<4)) + (n&7)*128 + (n>>3)*1024.
template <int R, bool PRECISE>
__device__ __forceinline__ void ln_store(const float (&h)[72], uint8_t* sm, uint32_t u_s, uint32_t pre_off, float alpha,
                                         float beta) {
  const float2* stats = reinterpret_cast<const float2*>(sm + OFF_STATS);
  static_assert(R % 2 == 0, "rows are stored in pairs");
#pragma unroll
  for (int n = 0; n < R; n += 2) {
    const float2 s0 = stats[n], s1 = stats[n + 1];
    const float u0 = fmaf((h[n] - s0.x) * s0.y, alpha, beta);
    const float u1 = fmaf((h[n + 1] - s1.x) * s1.y, alpha, beta);
    store_split2<PRECISE>(u_s, row_off(pre_off, n), row_off(pre_off, n + 1), u0, u1);
  }
}

// Walk the accumulator columns [0, R) of this thread's TMEM lane in chunks of 16 (+8) and hand
// (row, value) to f with compile-time row indices.  Must be executed by whole warps.
template <int R, class F>
__device__ __forceinline__ void for_acc(uint32_t taddr, F&& f) {
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
    tmem_ld16(taddr + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < R) f(c0 + j, v[j]);
  }
  {
    float v[8];
    tmem_ld8(taddr + 64, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (64 + j < R) f(64 + j, v[j]);
  }
}

// Token-mix bias through the GEMM: when the 72-row tile has a spare row (TED: 2S = 70), that row of
// U holds 1.0 and column 2S of the block-diagonal weight holds the bias, so the epilogue needs no
// per-row bias load.  BEAT (2S = 72) has no spare row and adds the bias from shared memory.
template <int S>
struct TokBias { static constexpr bool kInGemm = (2 * S + 1 <= 72); };

template <int S, bool PRECISE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT_ALL, 1) fused_step_kernel(const FusedParams p) {
  constexpr int R = 2 * S;            // real rows of a tile
  constexpr int NPRE = S - LS_F;      // prefix tokens per pass
  static_assert(R <= 72, "tile rows");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup ----------------------------------------------------------------
  for (uint32_t i = tid * 16; i < OFF_RING; i += NT_ALL * 16) *reinterpret_cast<uint4*>(sm + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&bars[BAR_FULL0 + s], 1);
      mbar_init(&bars[BAR_EMPTY0 + s], LS_MULTICAST ? 2 : 1);   // multicast: the MMA issuers of BOTH CTAs release a slot
    }
    for (int m = 0; m < 4; ++m) {
      mbar_init(&bars[BAR_UREADY0 + m], 128);     // the 4 warps that own the channels of M-tile m
      mbar_init(&bars[BAR_ACC0 + m], 1);
    }
    mbar_fence_init();
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  __syncthreads();
  if (TokBias<S>::kInGemm && tid < LS_D)          // the ones row (hi = 1.0, lo = 0) - never overwritten
    sts_u16(smem_u32(sm + OFF_UHI) + tile_off(R, tid, CBS), (uint16_t)0x3F80u);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster_sync_all();             // barriers of both CTAs initialised before any multicast touches them
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // The two CTAs of a cluster consume ONE weight stream (every stage is fetched half by each CTA and
  // multicast to both), so they run the same number of rounds; a CTA without a real clip in the
  // last round recomputes clip B-1 and writes nothing.
  const int n_rounds = (p.B + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t cta_rank = cluster_ctarank();

  if (warp >= 16) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
   // Both single-thread roles address shared memory with 32-bit shared-space addresses only
   // (32-register budget after setmaxnreg.dec).
   const uint32_t bars_s = smem_u32(bars), ring_s = smem_u32(sm + OFF_RING);
   const int n_layers = p.n_layers, KIN = p.KIN, MH = p.MH;
   if (warp == 16) {
    // ================= weight producers ==================================================
    // One thread sustains only ~1 bulk copy per 600 cycles whatever its size (bulk_bench.cu), so
    // four lanes issue in parallel: lane j owns ring slot j, i.e. every 4th stage of the stream.
    if (lane < NSLOT) {
      const uint8_t* const tape = p.tape;
      // stages consumed per round, in order: input projection | per layer: token-mix blocks,
      // channel-mix blocks | head.  PRECISE walks every tape stage; the bf16 mode skips the lo images.
      const uint32_t q_in = (PRECISE ? 8u : 4u) * (uint32_t)KIN, q_layer = PRECISE ? 68u : 34u;
      const uint32_t q_round = q_in + q_layer * (uint32_t)n_layers + (PRECISE ? 16u : 8u) * (uint32_t)MH;
      const uint32_t total = q_round * (uint32_t)n_rounds;
      const uint32_t full_s = bars_s + 8 * (BAR_FULL0 + lane), empty_s = bars_s + 8 * (BAR_EMPTY0 + lane);
      const uint32_t dst_s = ring_s + lane * SLOT;
      for (uint32_t it = lane; it < total; it += NSLOT) {
        const uint32_t q = it % q_round;
        uint32_t stage, bytes = W_HALF;
        if (PRECISE) {
          stage = q;
          if (q >= q_in && q < q_in + q_layer * (uint32_t)n_layers && (q - q_in) % 68u < 4u) bytes = WBLK_BLK;
        } else if (q < q_in) {
          stage = 2 * q;
        } else if (q < q_in + q_layer * (uint32_t)n_layers) {
          const uint32_t l = (q - q_in) / 34u, r = (q - q_in) % 34u;
          stage = 8u * (uint32_t)KIN + 68u * l + (r < 2 ? r : 4u + 2u * (r - 2u));
          if (r < 2) bytes = WBLK_BLK;
        } else {
          stage = 8u * (uint32_t)KIN + 68u * (uint32_t)n_layers + 2u * (q - q_in - 34u * (uint32_t)n_layers);
        }
        mbar_wait_s(empty_s, ((it / NSLOT) & 1) ^ 1);                   // (both CTAs are) done with the slot
        mbar_arrive_expect_tx_s(full_s, bytes);
#if LS_MULTICAST
        const uint32_t half = bytes >> 1;
        bulk_g2s_mc_s(dst_s + cta_rank * half, tape + (size_t)stage * SLOT + cta_rank * half, half, full_s, (uint16_t)3);
#else
        bulk_g2s_s(dst_s, tape + (size_t)stage * SLOT, bytes, full_s);
#endif
      }
    }
   } else if (warp == 17) {
    // ================= MMA issuer ==========================================================
    {   // all 32 lanes run this role in lock step; one elected lane issues (see ls_tc.cuh)
      uint32_t it = 0, uphase = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);    // provably warp-uniform copy
      constexpr uint32_t id_kk = idesc_bf16(128, NROW, 0, 0), id_mk = idesc_bf16(128, NROW, 1, 0);
      // Descriptor words: high word constant per layout, low word = (addr >> 4) | LBO field.
      constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
      const uint32_t u_s = smem_u32(sm + OFF_UHI);
      const uint32_t uk = desc_lo32(u_s, 16);          // K-major view of U_hi   (U_lo = + U_BYTES/16)
      const uint32_t um = desc_lo32(u_s, CBS);         // MN-major view of U_hi
      constexpr uint32_t LO = U_BYTES >> 4;            // descriptor-address distance U_hi -> U_lo
      auto wait_u_all = [&]() {       // whole operand tile published (K = all channels)
#pragma unroll
        for (int m = 0; m < 4; ++m) mbar_wait_s(bars_s + 8 * (BAR_UREADY0 + m), uphase & 1);
        ++uphase;
        tc_fence_after_sync();
      };
      auto wait_stage = [&]() -> uint32_t {       // -> descriptor low word of the next ring slot
        const uint32_t slot = it % NSLOT, ph = (it / NSLOT) & 1;
        mbar_wait_s(bars_s + 8 * (BAR_FULL0 + slot), ph);
        tc_fence_after_sync();
        return desc_lo32(ring_s + slot * SLOT, 16);
      };
      auto release_stage = [&]() {
#if LS_MULTICAST
        umma_commit_mc_s_elect(bars_s + 8 * (BAR_EMPTY0 + (it % NSLOT)), (uint16_t)3);
#else
        umma_commit_s_elect(bars_s + 8 * (BAR_EMPTY0 + (it % NSLOT)));
#endif
        ++it;
      };
      // D[mt] (+)= W[128 x 64 block] * U[:, 64*kc .. +64]^T   (weights = A, K-major; U = B, K-major)
      auto gemm_pair = [&](uint32_t d, uint32_t ub, bool first) {
        uint32_t wl = wait_stage();
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          umma_bf16_split_elect(d, wl + 2 * ks, DH, ub + 2 * ks, DH, id_kk, (first && ks == 0) ? 0u : 1u);
          if (PRECISE) umma_bf16_split_elect(d, wl + 2 * ks, DH, ub + LO + 2 * ks, DH, id_kk, 1u);
        }
        release_stage();
        if (PRECISE) {
          wl = wait_stage();
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) umma_bf16_split_elect(d, wl + 2 * ks, DH, ub + 2 * ks, DH, id_kk, 1u);
          release_stage();
        }
      };
      // one full GEMM over all of U: n_mt M-tiles x n_kc 64-channel blocks
      auto gemm_all = [&](int n_mt, int n_kc) {
        wait_u_all();
#pragma unroll 1
        for (int mt = 0; mt < 4; ++mt) {
          if (mt < n_mt)
#pragma unroll 1
            for (int kc = 0; kc < n_kc; ++kc) gemm_pair(tmem_u + (uint32_t)mt * NROW, uk + (uint32_t)kc * (CBS >> 4), kc == 0);
          umma_commit_s_elect(bars_s + 8 * (BAR_ACC0 + mt));    // M-tiles without work still flip their barrier
        }
      };
#pragma unroll 1
      for (int round = 0; round < n_rounds; ++round) {
        gemm_all(4, KIN);                                  // input projection
#pragma unroll 1
        for (int l = 0; l < n_layers; ++l) {
          // token mix: D[mt][ch, row_out] = sum_row_in U^T[ch, row_in] * Wblk[row_out, row_in].
          // Only the channels of M-tile mt are needed, so each M-tile starts as soon as ITS four
          // epilogue warps have published (per-M-tile barrier) and overlaps the others' LayerNorm.
          constexpr int NW = PRECISE ? 4 : 2;
          uint32_t wl[NW];
#pragma unroll
          for (int i = 0; i < NW; ++i) {
            wl[i] = wait_stage();
            ++it;
          }
          it -= NW;
#pragma unroll 1
          for (uint32_t mt = 0; mt < 4; ++mt) {
            mbar_wait_s(bars_s + 8 * (BAR_UREADY0 + mt), uphase & 1);
            tc_fence_after_sync();
            const uint32_t d = tmem_u + mt * NROW;
            const uint32_t ua = um + 2 * mt * (CBS >> 4);
#pragma unroll
            for (uint32_t ks = 0; ks < 5; ++ks) {
              const uint32_t bw = wl[ks >> 2] + 2 * (ks & 3), ao = ua + ks * (2048 >> 4);
              umma_bf16_split_elect(d, ao, DH, bw, DH, id_mk, ks == 0 ? 0u : 1u);
              if (PRECISE) {
                umma_bf16_split_elect(d, ao + LO, DH, bw, DH, id_mk, 1u);
                umma_bf16_split_elect(d, ao, DH, wl[NW - 2 + (ks >> 2)] + 2 * (ks & 3), DH, id_mk, 1u);
              }
            }
            umma_commit_s_elect(bars_s + 8 * (BAR_ACC0 + mt));
          }
          ++uphase;
#pragma unroll
          for (int i = 0; i < NW; ++i) release_stage();
          gemm_all(4, 8);                                  // channel mix
        }
        gemm_all(MH, 8);                                   // output head
      }
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
    // ================= epilogue: thread == channel ==========================================
    const int c = tid;
    const int mt = warp >> 2;
    const uint32_t lane_taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)mt * NROW;
    const uint32_t pre_off = (uint32_t)(c >> 6) * CBS + (uint32_t)(((c & 63) >> 3) << 4) + (uint32_t)(c & 7) * 2u;
    float* btok_s = reinterpret_cast<float*>(sm + OFF_BTOK);
    const uint32_t u_s = smem_u32(sm + OFF_UHI);
    const float emb = p.w.emb_table[(size_t)p.sp.t_model * LS_D + c];
    uint32_t aphase = 0;
    auto wait_acc = [&]() {
      mbar_wait(&bars[BAR_ACC0 + mt], aphase & 1);
      ++aphase;
      __syncwarp();                 // the spin loop may leave the warp diverged; tcgen05.ld is .aligned
      tc_fence_after_sync();
    };
    auto publish_u = [&]() {        // operand tile written (and accumulators consumed)
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&bars[BAR_UREADY0 + mt]);
    };
    float h[72];
    int tix = 0;
    auto stamp = [&]() {
      if (p.timing != nullptr && blockIdx.x == 0 && (tid == 0 || tid == 511) && tix < 256)
        p.timing[(tid ? 256 : 0) + tix++] = clock64();
    };

    for (int round = 0; round < n_rounds; ++round) {
      const int b_raw = (int)blockIdx.x + round * (int)gridDim.x;
      const bool valid = b_raw < p.B;
      const int b = valid ? b_raw : p.B - 1;
      // ---- X operand: x_t[b] (hi, lo) into rows p*S + NPRE + f, k = j ----------------------
      const float* xb = p.x_t + (size_t)b * p.JD * LS_F;
      for (int i = tid; i < p.JD * LS_F; i += NT_EPI) {
        const int j = i / LS_F, f = i - j * LS_F;
        const float v = xb[i];
        store_split<PRECISE>(u_s, tile_off(NPRE + f, j, CBS), v);
        store_split<PRECISE>(u_s, tile_off(S + NPRE + f, j, CBS), v);
      }
      publish_u();
      stamp();   // 0: X operand published
      // ---- residual stream init: hoisted terms now, projection result when it lands --------
      {
        const float mu = p.z_mu[(size_t)b * LS_D + c], sd = __expf(0.5f * p.z_lv[(size_t)b * LS_D + c]);
        h[0] = fmaf(p.eps_c[(size_t)b * LS_D + c], sd, mu);
        h[S] = fmaf(p.eps_u[(size_t)b * LS_D + c], sd, mu);
        if (NPRE == 2) h[1] = h[S + 1] = p.emo_tok[(size_t)b * LS_D + c];
        const float* Pb = p.P + (size_t)b * LS_F * LS_D + c;
        const float* Ab = p.A + (size_t)b * LS_F * LS_D + c;
#pragma unroll
        for (int f = 0; f < LS_F; ++f) {
          const float pv = Pb[f * LS_D];
          h[S + NPRE + f] = pv;
          h[NPRE + f] = pv + Ab[f * LS_D];
        }
      }
      wait_acc();
      for_acc<R>(lane_taddr, [&](int n, float v) {
        if ((n % S) >= NPRE) h[n] += v;          // prefix-token rows keep their direct values
      });
      stamp();   // 1: input projection consumed

      for (int l = 0; l < p.n_layers; ++l) {
        const LsLayerW L = p.w.layer[l];
        const float a1 = L.ln1_a[c], b1 = L.ln1_b[c], a2 = L.ln2_a[c], b2 = L.ln2_b[c], bch = L.b_ch[c];
        if (tid < S) btok_s[tid] = btok_s[S + tid] = L.b_tok[tid];
        // x = x + emb ; LN1 ; -> operand tile
#pragma unroll
        for (int n = 0; n < R; ++n) h[n] += emb;
        if (l == 0) ln_stats<R>(h, sm, false);     // provisional means for the shift
        ln_stats<R>(h, sm, true);
        stamp();   // LN1 stats done
        ln_store<R, PRECISE>(h, sm, u_s, pre_off, a1, b1);
        publish_u();
        stamp();   // U1 published
        // token mix epilogue: x = x + silu(conv + bias)
        wait_acc();
        stamp();   // token-mix accumulator ready
        for_acc<R>(lane_taddr, [&](int n, float v) { h[n] += silu_fast(TokBias<S>::kInGemm ? v : v + btok_s[n]); });
        stamp();   // token-mix epilogue done
        ln_stats<R>(h, sm, true);
        stamp();   // LN2 stats done
        ln_store<R, PRECISE>(h, sm, u_s, pre_off, a2, b2);
        publish_u();
        stamp();   // U2 published
        // channel mix epilogue: x = x + silu(linear + bias)
        wait_acc();
        stamp();   // channel-mix accumulator ready
        for_acc<R>(lane_taddr, [&](int n, float v) { h[n] += silu_fast(v + bch); });
        stamp();   // channel-mix epilogue done
      }

      // ---- output head operand: plain hi/lo split of h -------------------------------------
      // Every thread must be past its last wait_acc before U is overwritten: the channel-mix
      // MMAs of the other M-tiles still read U until their own accumulator barrier fires.
      epi_bar();
#pragma unroll
      for (int n = 0; n < R; n += 2)
        store_split2<PRECISE>(u_s, row_off(pre_off, n), row_off(pre_off, n + 1), h[n], h[n + 1]);
      publish_u();
      wait_acc();
      if (warp * 32 < p.JD) {                    // warp-uniform: tcgen05.ld is .aligned
        // h is dead: reuse it for the head outputs (lane = output feature j, column = row)
        for_acc<R>(lane_taddr, [&](int n, float v) { h[n] = v; });
        if (c < p.JD && valid) {
          const float bo = p.w.b_out[c], sc = p.scale[b];
          const size_t base = ((size_t)b * p.JD + c) * LS_F;
          const float* nzp = p.noise ? p.noise + (size_t)b * p.nsb + (size_t)c * p.nsj : nullptr;
          const bool use_noise = (p.sp.mode != 2) && p.sp.add_noise && nzp != nullptr;
#pragma unroll
          for (int f = 0; f < LS_F; ++f) {
            const float oc = h[NPRE + f] + bo, ou = h[S + NPRE + f] + bo;
            float x0 = ou + sc * (oc - ou);                       // cfg_sampler.py:31
            const float nz = use_noise ? nzp[(size_t)f * p.nsf] : 0.f;
            float xp = 0.f;
            x0 = ls_sampler_update(p.sp, x0, p.x_t[base + f], nz, &xp);
            if (p.pred_x0) p.pred_x0[base + f] = x0;
            if (p.sp.mode != 2) p.x_prev[base + f] = xp;
          }
        }
      }
      // the next tile's X operand overwrites U: every head MMA has completed (wait_acc above)
      tc_fence_before_sync();
      epi_bar();
    }
  }
  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before_sync();
  cluster_sync_all();             // the peer may still multicast commits into this CTA's barriers
  if (warp == 17) tmem_dealloc<512>(tmem);
}

// ---- weight tape ------------------------------------------------------------------------------
// stage = [128 x 64] block of W (rows m0.., cols k0..) as K-major swizzle-128B images, hi then lo.
__global__ void build_w_stage_kernel(const float* __restrict__ src, int rows, int cols, int ld, int m0, int k0,
                                     uint8_t* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 128 * 64; i += gridDim.x * blockDim.x) {
    const int m = i >> 6, k = i & 63;
    const float v = (m0 + m < rows && k0 + k < cols) ? src[(size_t)(m0 + m) * ld + k0 + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const uint32_t off = tile_off(m, k, 0);
    *reinterpret_cast<__nv_bfloat16*>(dst + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(dst + W_HALF + off) = lo;
  }
}

// token-mix stages: block-diagonal [80 x 128] (two passes share W_tok), K-major; one stage per
// 64-wide k block: hi(k<64), hi(k>=64), lo(k<64), lo(k>=64)
__global__ void build_wblk_kernel(const float* __restrict__ w_tok, const float* __restrict__ b_tok, int S,
                                  uint8_t* __restrict__ dst_hi, uint8_t* __restrict__ dst_lo) {
  const bool bias_col = (2 * S + 1 <= 72);         // must match TokBias<S>::kInGemm
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NROW * 128; i += gridDim.x * blockDim.x) {
    const int n = i >> 7, k = i & 127;
    float v = 0.f;
    if (n < 2 * S && k < 2 * S && (n / S) == (k / S)) v = w_tok[(n % S) * S + (k % S)];
    if (bias_col && n < 2 * S && k == 2 * S) v = b_tok[n % S];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const uint32_t off = (uint32_t)(k >> 6) * SLOT + tile_off(n, k & 63, 0);
    *reinterpret_cast<__nv_bfloat16*>(dst_hi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(dst_lo + off) = lo;
  }
}

struct FusedState {
  uint8_t* tape = nullptr;
  size_t tape_bytes = 0;
  int KIN = 0, MH = 0, n_stages = 0;
  int sm_count = 0;
  bool attr_done[2][2] = {{false, false}, {false, false}};
};

template <int S, bool PRECISE>
int launch_fused(ls_handle* h, FusedState* fs, const FusedParams& fp, cudaStream_t s) {
  bool& done = fs->attr_done[S - 35][PRECISE ? 1 : 0];
  if (!done) {
    LS_CUDA(h, cudaFuncSetAttribute(fused_step_kernel<S, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DYN));
    done = true;
  }
  // clusters of 2: an even grid, at most one CTA per SM
  const int grid = (fp.B + 1 < fs->sm_count ? fp.B + 1 : fs->sm_count) & ~1;
  fused_step_kernel<S, PRECISE><<<grid, NT_ALL, SMEM_DYN, s>>>(fp);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

}  // namespace

int lsf_available(const ls_handle* h) { return h && h->fused != nullptr; }

void lsf_destroy(ls_handle* h) {
  if (h && h->fused) {
    FusedState* fs = static_cast<FusedState*>(h->fused);
    if (fs->tape) cudaFree(fs->tape);
    delete fs;
    h->fused = nullptr;
  }
}

// Builds the weight tape.  Returns 0 when the fused path is usable, 1 when it is not built
// for this geometry, < 0 on error.
int lsf_init(ls_handle* h, cudaStream_t s) {
  if (h->S != 35 && h->S != 36) return 1;
  FusedState* fs = static_cast<FusedState*>(h->fused);
  if (!fs) {
    fs = new FusedState();
    fs->KIN = (h->JD + 63) / 64;
    fs->MH = (h->JD + 127) / 128;
    if (fs->KIN > 8 || fs->MH > 4) {
      delete fs;
      return 1;
    }
    fs->n_stages = 8 * fs->KIN + h->cfg.n_layers * 68 + 16 * fs->MH;
    fs->tape_bytes = (size_t)fs->n_stages * SLOT;
    if (cudaMalloc(&fs->tape, fs->tape_bytes) != cudaSuccess) {
      delete fs;
      return ls_fail(h, LS_ENOMEM, "weight tape (%zu bytes)", (size_t)fs->n_stages * SLOT);
    }
    cudaDeviceGetAttribute(&fs->sm_count, cudaDevAttrMultiProcessorCount, h->cfg.device);
    h->fused = fs;
  }
  LS_CUDA(h, cudaMemsetAsync(fs->tape, 0, fs->tape_bytes, s));
  auto raw = [&](const std::string& k) -> const float* {
    for (auto& r : h->raw)
      if (r.key == k) return r.dev;
    return nullptr;
  };
  const int IN = 2 * h->JD + 1 + LS_AF;
  size_t st = 0;
  const float* win = raw("input_mapping.weight");
  for (int mt = 0; mt < 4; ++mt)
    for (int kc = 0; kc < fs->KIN; ++kc) {
      build_w_stage_kernel<<<16, 256, 0, s>>>(win, LS_D, h->JD, IN, mt * 128, kc * 64, fs->tape + st * SLOT);
      LS_LAUNCH_CHECK(h);
      st += 2;
    }
  for (int l = 0; l < h->cfg.n_layers; ++l) {
    const std::string p = "backbone.mlps." + std::to_string(l) + ".";
    build_wblk_kernel<<<16, 256, 0, s>>>(raw(p + "block1.1.weight"), raw(p + "block1.1.bias"), h->S, fs->tape + st * SLOT,
                                         fs->tape + (st + 2) * SLOT);
    LS_LAUNCH_CHECK(h);
    st += 4;
    const float* wch = raw(p + "block2.1.weight");
    for (int mt = 0; mt < 4; ++mt)
      for (int kc = 0; kc < 8; ++kc) {
        build_w_stage_kernel<<<16, 256, 0, s>>>(wch, LS_D, LS_D, LS_D, mt * 128, kc * 64, fs->tape + st * SLOT);
        LS_LAUNCH_CHECK(h);
        st += 2;
      }
  }
  const float* wout = raw("output_process.poseFinal.weight");
  for (int mt = 0; mt < fs->MH; ++mt)
    for (int kc = 0; kc < 8; ++kc) {
      build_w_stage_kernel<<<16, 256, 0, s>>>(wout, h->JD, LS_D, LS_D, mt * 128, kc * 64, fs->tape + st * SLOT);
      LS_LAUNCH_CHECK(h);
      st += 2;
    }
  if ((int)st != fs->n_stages) return ls_fail(h, LS_EINVAL, "tape stage count mismatch");
  return LS_OK;
}

int lsf_step(ls_handle* h, int B, const ls_step_params* p, int precise, const float* x_t, const float* eps_c,
             const float* eps_u, const float* noise, int64_t sb, int64_t sj, int64_t sf, const float* scale,
             float* x_prev, float* pred_x0, cudaStream_t s) {
  FusedState* fs = static_cast<FusedState*>(h->fused);
  if (!fs) return ls_fail(h, LS_EUNSUPPORTED, "tcgen05 path not available");
  if (x_prev == x_t && p->mode != 2)
    ;  // in-place update is fine: each element is read before it is written by the same thread
  FusedParams fp{};
  fp.tape = fs->tape;
  fp.n_layers = h->cfg.n_layers;
  fp.JD = h->JD;
  fp.KIN = fs->KIN;
  fp.MH = fs->MH;
  fp.B = B;
  fp.w = h->w;
  fp.A = h->A; fp.P = h->P; fp.z_mu = h->z_mu; fp.z_lv = h->z_lv; fp.emo_tok = h->emo_tok;
  fp.x_t = x_t; fp.eps_c = eps_c; fp.eps_u = eps_u; fp.noise = noise; fp.scale = scale;
  fp.nsb = sb; fp.nsj = sj; fp.nsf = sf;
  fp.x_prev = x_prev; fp.pred_x0 = pred_x0;
  fp.sp = *p;
  fp.timing = nullptr;
  static long long* tbuf = nullptr;
  const bool timing = getenv("LS_FUSED_TIMING") != nullptr;
  if (timing) {
    if (!tbuf) cudaMalloc(&tbuf, 512 * sizeof(long long));
    cudaMemsetAsync(tbuf, 0, 512 * sizeof(long long), s);
    fp.timing = tbuf;
  }
  int rc;
  if (h->S == 35) rc = precise ? launch_fused<35, true>(h, fs, fp, s) : launch_fused<35, false>(h, fs, fp, s);
  else rc = precise ? launch_fused<36, true>(h, fs, fp, s) : launch_fused<36, false>(h, fs, fp, s);
  if (timing && rc == LS_OK) {
    long long t[512];
    cudaStreamSynchronize(s);
    cudaMemcpy(t, tbuf, sizeof(t), cudaMemcpyDeviceToHost);
    static const char* names[] = {"LN1 stats", "U1 publish", "tok acc wait", "tok epilogue", "LN2 stats", "U2 publish",
                                  "ch acc wait", "ch epilogue"};
    for (int w = 0; w < 2; ++w) {
      const long long* q = t + 256 * w;
      fprintf(stderr, "[fused timing] thread %d: X publish -> in-proj consumed %lld cyc\n", w ? 511 : 0, q[1] - q[0]);
      for (int l = 0; l < h->cfg.n_layers && 2 + 8 * l + 7 < 256; ++l) {
        fprintf(stderr, "  layer %d:", l);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %s %lld |", names[k], q[2 + 8 * l + k] - q[1 + 8 * l + k]);
        fprintf(stderr, "\n");
      }
    }
  }
  return rc;
}
