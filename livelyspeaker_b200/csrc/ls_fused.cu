// The fused denoising step(s): ONE persistent sm_100a kernel per launch of up to 16 steps.
//
//   _WrappedModel + ClassifierFreeSampleModel + RAG.forward (minus the hoisted terms) +
//   p_mean_variance + p_sample / ddim_sample of the reference
//   (respace.py:118-130, cfg_sampler.py:24-31, RAG.py:98-133, mlp_module.py:37-91,
//    gaussian_diffusion.py:284-399, 507-558, 745-798) - see DESIGN.md "fused step kernel".
//
// Work item (tile) = one clip at one step, both guidance passes: 2*S token rows (S = 35 TED / 36 BEAT).
// A launch covers n_steps consecutive steps: item q = step * B + clip, CTA i takes q = i, i+grid, ...
// Item (k, b) needs x of (k-1, b), produced >= 3 rounds earlier by another CTA and published through a
// per-clip release/acquire counter, so the 512 clips of a batch no longer quantise to 4 rounds of 148.
//
// All GEMMs run TRANSPOSED on the tensor cores, D^T[channel, row] = W[channel, k] * U^T[k, row]:
//   M = 128 output channels (4 M-tiles for d = 512), N = rows of the tile, K = channels,
// so the TMEM lane of an accumulator element is its CHANNEL and the column is its ROW:
//   * each of the 512 epilogue threads keeps 4 channels x 18 rows of the fp32 residual stream
//     in REGISTERS for the whole 8-layer stack;
//   * TMEM holds only accumulators;
//   * shared memory holds the bf16 operand tile U of the channel-type GEMMs (input projection, channel mix,
//     head) as their K-major B operand, plus a 4-slot ring of weight stages streamed from L2 by 1-D bulk async
//     copies; the token mix takes its A operand (LayerNorm 1's output) from TENSOR MEMORY, written there by the
//     epilogue threads (TS mode).
// Precision: PRECISE = bf16x3 (hi*hi + hi*lo + lo*hi, fp32 accumulate) meets the rtol 1e-3 /
// atol 1e-4 parity bar; !PRECISE = plain bf16 operands (fast mode).
//
// What bounds a GEMM phase is SHARED-MEMORY BANDWIDTH, not the tensor pipe: an SS-mode M=128,K=16 MMA
// costs max(N/2, 32 + N/4) cycles (umma_bench2.cu: the 4 KB A block and the N*32 B of B are read at
// 128 B/clk), and the weight ring fill shares that bandwidth.  Hence the operand tile stores, per
// 64-channel block, the lo image (72 rows) directly followed by the hi image (72 rows): ONE N=144 MMA
// reads a W_hi block once for both U_lo and U_hi (columns [0,72) and [72,144) of the accumulator),
// a second N=80 MMA adds W_lo * U_hi into columns [72,152); the epilogue sums the two column ranges.
//
// LayerNorm 2 is off the critical path: the channel-mix operand is normalised with the statistics of
// LayerNorm 1 of the same rows (already known), alpha2 is folded into the weight tape and the exact
// statistics - reduced while the GEMM runs - enter as a per-row scale / shift in the GEMM epilogue:
//   W.(alpha*((h-mu)*rho)+beta)+b = (rho/s)*acc + ((m-mu)*rho)*S_c + t_c,  acc = (W.alpha)*((h-m)*s).
//
// Warp roles (20 warps): 0-15 epilogue, 16 weight producer, 17 MMA issuer + TMEM owner, 18-19 idle; warps
// 16-19 give most of their registers to the epilogue warps with setmaxnreg (112 vs 32) so the residual stream fits.
// Epilogue thread (lq = warp & 3, rq = warp >> 2, lane) owns TMEM lane 32 lq + lane of every M-tile (four
// channels) and 18 of the 72 tile rows: see "epilogue thread mapping" below.
#include <cuda_bf16.h>
#include <cstdlib>

#include "ls_internal.cuh"
#include "ls_tc.cuh"
#include "ls_update.cuh"

using namespace lstc;

namespace {

constexpr int NT_EPI = 512;
// Register file: 16384 registers per SM sub-partition, 5 warps on each (20 warps).  Launch at 96 registers per
// thread (5*32*96 = 15360 per sub-partition); setmaxnreg then moves them to the 4 epilogue warps of the
// sub-partition: 4*32*112 + 32*32 == 15360.  18 warps at 112 do not launch ("too many resources": a
// sub-partition would hold 5 warps * 3584), and a larger total blocks forever in setmaxnreg.inc.
constexpr int NT_ALL = 640;                     // 16 epilogue warps + 1 producer + 1 MMA + 2 idle (register donors)
constexpr int REGS_EPI = 112;
constexpr int REGS_AUX = 32;
constexpr int NROW = 80;                        // MMA N of a single-image product (rows of a tile, padded)
constexpr int NCAT = 144;                       // MMA N of the concatenated [U_lo ; U_hi] operand
constexpr int RGS = 9;                          // 8-row groups per image (72 rows)
constexpr uint32_t HALF_BLK = RGS * 1024;       // one image (lo or hi) of a 64-channel block
constexpr uint32_t CBS = 2 * HALF_BLK;          // bytes between 64-channel blocks of U: [lo image][hi image]
constexpr uint32_t HI_OFF = HALF_BLK;           // hi image inside a block
constexpr uint32_t U_BYTES = 8 * CBS;
constexpr uint32_t SLOT = 16384;                // ring slot = tape stage pitch
constexpr int NSLOT = 4;
#ifndef LS_MULTICAST
// 1: CTAs run as cluster pairs, every weight stage is fetched half by each CTA and multicast to both (half the
//    L2 reads, but the pair advances in lock step through the shared ring).  0 (default): independent CTAs, every
//    CTA streams the whole tape from L2 - measured 2 % faster at B = 512 (1255 vs 1229 steps/s): L2 delivers the
//    5.9 TB/s without trouble and the shared-memory pipe, not L2, bounds the GEMM phases.
#define LS_MULTICAST 0
#endif
#if LS_MULTICAST
#define LS_CLUSTER_ATTR __cluster_dims__(2, 1, 1)
#else
#define LS_CLUSTER_ATTR
#endif
#ifndef LS_LN1_EARLY
// 1: the channel-mix epilogue of block l adds block l+1's time embedding and accumulates its LayerNorm-1 partial sums
//    M-tile by M-tile (ln_partial_q), so that only the last butterfly and the 4-warp combine follow the last MMA.
//    Measured (profiles/r2_ab_power_cap.txt): the best launch gets 0.5 % shorter, but the kernel sits at the 1000 W power
//    cap and the extra work (+ 15 % warp instructions, + 33 % LSU shared wavefronts: three more butterflies per block)
//    comes back as a 0.8 % lower clock - net + 0.1 %; + 1.0 % only in the uncapped multicast build.  0 (default): off.
#define LS_LN1_EARLY 0
#endif
#ifndef LS_NOFETCH
#define LS_NOFETCH 0     // diagnostic: 1 = the producer signals stages without copying (garbage results, pure MMA timing)
#endif
#ifndef LS_MMA_PROF
#define LS_MMA_PROF 0    // diagnostic: 1 = the MMA warp accounts its cycles (ring waits / operand waits / total) of CTA 0
#endif
constexpr uint32_t W_HALF = 16384;              // 128 x 64 bf16 weight block (hi or lo) = one stage
constexpr uint32_t WBLK_BLK = 10 * 1024;        // token-mix weights: 80 rows x 64 k = one stage
// TMEM columns.  Channel-type GEMMs (input projection, channel mix) alternate between two buffers of ACC_COLS:
// [0,72) W_hi*U_lo, [72,144) W_hi*U_hi + W_lo*U_hi, [144,152) overhang of the N=80 product.  The token mix
// alternates between two N=80 buffers behind them, so a channel mix may start while the last token-mix
// accumulators are still being read.  The head (<= 3 M-tiles, read only when all are complete) uses three
// ACC_COLS buffers from column 0.
constexpr int ACC_COLS = 152;
constexpr int NACC = 3;                         // head M-tiles at most
constexpr int CAT_HI = 72;                      // first column of the hi-image product in a buffer
// Token mix (TS mode): the LayerNorm-1 output goes from the epilogue threads straight into tensor memory as the
// K-major A operand (hi image 40 columns = 80 rows packed in pairs, lo image 40 columns): two alternating operand
// buffers and two alternating N=80 accumulators.  The accumulators alias channel buffer 1 (a token-mix MMA is
// issued only after all 16 warps have published its operand, i.e. finished reading the previous channel mix; the
// channel mix reaches buffer 1 only after every token-mix accumulator was read), the operand buffers alias
// nothing that another warp still reads.  Channel buffer 0 stays free for the early start of the channel mix.
constexpr int TOK_D0 = ACC_COLS;                // accumulators:    TOK_D0 + (m & 1) * 80
constexpr int TOK_A0 = ACC_COLS + 2 * NROW;     // operand buffers: TOK_A0 + (m & 1) * 80  (hi [0,40), lo [40,80))
constexpr int TOK_AIMG = 40;
constexpr int KMAX = LS_MAX_FUSED_STEPS;

// shared memory map (offsets from a 1024-aligned base)
constexpr uint32_t OFF_U = 0;
constexpr uint32_t OFF_PAD = U_BYTES;                   // 1024 zero bytes: row group 9 of the last hi image
constexpr uint32_t OFF_RING = OFF_PAD + 1024;
constexpr uint32_t OFF_PART = OFF_RING + NSLOT * SLOT;  // [16 warps][18 rows] float2 LN partial sums
constexpr uint32_t OFF_STATS = OFF_PART + 16 * 18 * 8;  // [72] float2 (rstd, -mean * rstd)
constexpr uint32_t OFF_MEAN = OFF_STATS + 72 * 8;       // [72] float mean (the next statistics' shift)
constexpr uint32_t OFF_GD = OFF_MEAN + 72 * 4;          // [72] float2 channel-mix epilogue (row scale, row shift)
constexpr uint32_t OFF_BTOK = OFF_GD + 72 * 8;          // [72] float token-mix bias
constexpr uint32_t OFF_PAB = OFF_BTOK + 72 * 4;         // [512] float2 (ln1 alpha, beta) of the current layer, thread-private slots
constexpr uint32_t OFF_PSC = OFF_PAB + 512 * 8;         // [512] float2 (S_c, t_c) of the current layer, thread-private slots
constexpr uint32_t OFF_EMB = OFF_PSC + 512 * 8;         // [512] float time embedding of the current item, thread-private slots
constexpr uint32_t OFF_BARS = OFF_EMB + 512 * 4;         // mbarriers
constexpr uint32_t OFF_TMEM = OFF_BARS + 24 * 8;
constexpr uint32_t SMEM_USED = OFF_TMEM + 16;
constexpr uint32_t SMEM_DYN = SMEM_USED + 1024;         // + alignment slack
static_assert(SMEM_DYN <= 232448, "shared memory budget");
static_assert(NACC * ACC_COLS <= 512 && TOK_A0 + 2 * NROW <= 512, "TMEM budget");

enum { BAR_FULL0 = 0, BAR_EMPTY0 = 4, BAR_UREADY0 = 8, BAR_ACC0 = 12, BAR_DRAIN0 = 16, BAR_TDRAIN0 = 18 };   // indices into the mbarrier array

struct StepIO {
  const float* eps_c; const float* eps_u; const float* noise;
  long long nsb, nsj, nsf;
  float* x_prev; float* pred_x0;
};

struct FusedParams {
  const uint8_t* tape;     // weight stages, SLOT bytes apart
  int n_layers, JD, KIN, MH, B, n_steps;
  LsWeights w;
  const float* Sc;         // [n_layers][512] sum_k W_ch[c,k]*alpha2[k]
  const float* tc;         // [n_layers][512] sum_k W_ch[c,k]*beta2[k] + b_ch[c]
  const float* A; const float* P; const float* z_mu; const float* z_lv; const float* emo_tok;
  const float* x_in; const float* scale;
  int* flags;              // [B] steps completed per clip in this launch (n_steps > 1)
  const long long* t_dev;  // per-clip ORIGINAL timesteps on the device (ls_cfg_forward, mode 2) or nullptr: sp.t_model
  int max_t;               // rows of the time-embedding table (t_dev values are clamped into it)
  float* dbg_h;            // ls_debug_hidden: [B][2][S][512] residual stream after layer dbg_layer of step 0, or nullptr
  int dbg_layer;
  ls_step_params sp[KMAX];
  StepIO io[KMAX];
  long long* timing;       // debug (LS_FUSED_TIMING=1): clock64 stamps of block 0, threads 0 and 511
};

// x * sigmoid(x) with ex2.approx + rcp.approx (about 2 ulp; the IEEE reciprocal costs ~10 more instructions).
// The .ftz forms on purpose: __expf / __fdividef wrap the same two MUFU ops in range fix-ups (FSETP + two predicated
// FMULs per call for results below 2^-126, a select for huge divisors) that cannot matter here - e flushed to zero
// gives z / 1, e = +inf gives z * 0 - and the SiLU is 15 % of all instructions the epilogue warps issue (ncu, round 1):
// 5 instructions per call instead of 8.
__device__ __forceinline__ float silu_fast(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return z * r;
}

// (h0, h1) += silu(z0), silu(z1): FMUL2, 2 x EX2, FADD2, 2 x RCP, FFMA2 - 7 instructions for two activations
__device__ __forceinline__ void silu_acc2(float& h0, float& h1, f32x2 z) {
  float t0, t1, e0, e1, d0, d1, r0, r1;
  upk2(mul2(z, bc2(-1.4426950408889634f)), t0, t1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  upk2(add2(pk2(e0, e1), bc2(1.f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  upk2(fma2(z, pk2(r0, r1), pk2(h0, h1)), h0, h1);
}

// Per-row statistics live in shared memory in PAIR layout: rows (2i, 2i+1) share one float4 = (x_2i, x_2i+1, y_2i,
// y_2i+1), so that one 128-bit load hands a thread the packed operands of both rows of a register pair.
__device__ __forceinline__ int pair_slot(int row) { return ((row >> 1) << 2) + (row & 1); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// No "memory" clobber: the operand tile is written only through these stores and read only by the async proxy
// (after fence.proxy.async, itself a volatile asm that stays behind them), so the compiler may hoist the shared
// loads of the next row pair above the stores of this one instead of serialising pair after pair.
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(v));
}
// hi/lo bf16 split of v, stored at byte offset `off` of the lo image (hi image = + HI_OFF); u_s = shared-space address of U
template <bool PRECISE>
__device__ __forceinline__ void store_split(uint32_t u_s, uint32_t off, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  sts_u16(u_s + HI_OFF + off, __bfloat16_as_ushort(hi));
  if (PRECISE) {
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    sts_u16(u_s + off, __bfloat16_as_ushort(lo));
  }
}

// ---- epilogue thread mapping ---------------------------------------------------------------------
// Thread (lq = warp & 3, rq = warp >> 2, lane) owns TMEM lane 32*lq + lane of EVERY M-tile, i.e. the four
// channels c_m = 128*m + 32*lq + lane, and the NQ = 18 tile rows [18*rq, 18*rq + 18).  So
//   * every accumulator is drained by all 16 warps (18 columns each) as soon as its M-tile completes - the
//     tail after the last M-tile of a GEMM is a quarter of what a warp-per-M-tile mapping leaves;
//   * a LayerNorm row lives in ONE group of 4 warps (rq): 4 local adds, a 38-shuffle transposing butterfly,
//     a 4-way exchange through shared memory and two 128-thread named barriers - no CTA-wide barrier.
constexpr int NQ = 18;

__device__ __forceinline__ void group_bar(int rq) { asm volatile("bar.sync %0, 128;" ::"r"(2 + rq) : "memory"); }

// transposing butterfly: V values per lane in, the sum over all 32 lanes of value idx(lane) out in v[0]
template <int V>
__device__ __forceinline__ void butterfly(float* v, int lane) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int xm = 16 >> s, half = (V >> 1) >> s;
    if (half >= 1) {
      const bool up = (lane & xm) != 0;
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float send = up ? v[j] : v[j + half];
        const float keep = up ? v[j + half] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, xm);
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], xm);
    }
  }
}

template <int CORR>
__device__ __forceinline__ void ln_finalize_q(uint8_t* sm, int rq, bool use_shift);

// Per-row (sum, sum of squares) over the 512 channels of this group's 18 rows.  `shift` = the previous mean of
// the row (robust single-pass variance), or 0 when use_shift is false.
//   CORR == 0: stats[row] = (rstd, -mean * rstd) (normalisation is one FMA; pair layout), means[row] = mean.
//   CORR == 1: (m, s) = the mean / rstd the operand tile was normalised with; the exact (mu, rho)
//              give the channel-mix epilogue's gd[row] = (rho / s, (m - mu) * rho) and replace stats / means.
template <int CORR>
__device__ __forceinline__ void ln_stats_q(const float (&h)[72], uint8_t* sm, int rq, bool use_shift) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* part = reinterpret_cast<float*>(sm + OFF_PART);          // [16 warps][18 rows][2]
  float* stats = reinterpret_cast<float*>(sm + OFF_STATS);        // pair layout (rstd, -mean * rstd)
  float* means = reinterpret_cast<float*>(sm + OFF_MEAN);
  const int r0 = NQ * rq;
  // rows j, j+1 (a register pair of every M-tile) at once: packed (sum, sum of squares) over the 4 channels
  auto row_sq2 = [&](int j, float& s0, float& s1, float& q0, float& q1) {
    f32x2 d0 = pk2(h[j], h[j + 1]), d1 = pk2(h[NQ + j], h[NQ + j + 1]), d2 = pk2(h[2 * NQ + j], h[2 * NQ + j + 1]),
          d3 = pk2(h[3 * NQ + j], h[3 * NQ + j + 1]);
    if (use_shift) {
      const float2 m2 = *reinterpret_cast<const float2*>(means + r0 + j);
      const f32x2 sh = pk2(m2.x, m2.y);
      d0 = sub2(d0, sh);
      d1 = sub2(d1, sh);
      d2 = sub2(d2, sh);
      d3 = sub2(d3, sh);
    }
    upk2(add2(add2(d0, d1), add2(d2, d3)), s0, s1);
    upk2(add2(fma2(d0, d0, mul2(d1, d1)), fma2(d2, d2, mul2(d3, d3))), q0, q1);
  };
#pragma unroll
  for (int g = 0; g < 2; ++g) {          // rows 8g .. 8g+7: values 0..7 = sums, 8..15 = sums of squares
    float v[16];
#pragma unroll
    for (int i = 0; i < 8; i += 2) row_sq2(8 * g + i, v[i], v[i + 1], v[8 + i], v[9 + i]);
    butterfly<16>(v, lane);              // lane bits (4 | 3,2,1) select (q? | row)
    if ((lane & 1) == 0) part[((warp * NQ) + 8 * g + ((lane >> 1) & 7)) * 2 + (lane >> 4)] = v[0];
  }
  {                                      // rows 16, 17
    float v[4];
    row_sq2(16, v[0], v[1], v[2], v[3]);
    butterfly<4>(v, lane);               // lane bits (4 | 3) select (q? | row)
    if ((lane & 7) == 0) part[((warp * NQ) + 16 + ((lane >> 3) & 1)) * 2 + (lane >> 4)] = v[0];
  }
  ln_finalize_q<CORR>(sm, rq, use_shift);
}

// One M-tile's share of the NEXT LayerNorm-1 statistics, accumulated into this warp's own partial-sum slots while the
// tensor pipe still works on the later M-tiles of the channel mix: hm = the thread's 18 rows of channel c_m, already
// final (+ the next block's time embedding).  Shift = means[] (LayerNorm 2's exact mean of the same rows).  After the
// last M-tile only ln_finalize_q<0> is left on the critical path - the 4-channel sums, three of the four butterflies
// and their shuffle latency are off it.  No barrier: a lane only ever touches its own slots.
template <bool FIRST>
__device__ __forceinline__ void ln_partial_q(const float* hm, uint8_t* sm, int rq) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* part = reinterpret_cast<float*>(sm + OFF_PART);
  const float* means = reinterpret_cast<const float*>(sm + OFF_MEAN);
  const int r0 = NQ * rq;
  auto row2 = [&](int j, float& s0, float& s1, float& q0, float& q1) {
    const float2 m2 = *reinterpret_cast<const float2*>(means + r0 + j);
    const f32x2 d = sub2(pk2(hm[j], hm[j + 1]), pk2(m2.x, m2.y));
    upk2(d, s0, s1);
    upk2(mul2(d, d), q0, q1);
  };
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 8; i += 2) row2(8 * g + i, v[i], v[i + 1], v[8 + i], v[9 + i]);
    butterfly<16>(v, lane);
    if ((lane & 1) == 0) {
      float* slot = part + ((warp * NQ) + 8 * g + ((lane >> 1) & 7)) * 2 + (lane >> 4);
      *slot = FIRST ? v[0] : *slot + v[0];
    }
  }
  {
    float v[4];
    row2(16, v[0], v[1], v[2], v[3]);
    butterfly<4>(v, lane);
    if ((lane & 7) == 0) {
      float* slot = part + ((warp * NQ) + 16 + ((lane >> 3) & 1)) * 2 + (lane >> 4);
      *slot = FIRST ? v[0] : *slot + v[0];
    }
  }
}

template <int CORR>
__device__ __forceinline__ void ln_finalize_q(uint8_t* sm, int rq, bool use_shift) {
  const int tid = threadIdx.x;
  float* part = reinterpret_cast<float*>(sm + OFF_PART);          // [16 warps][18 rows][2]
  float* stats = reinterpret_cast<float*>(sm + OFF_STATS);        // pair layout (rstd, -mean * rstd)
  float* means = reinterpret_cast<float*>(sm + OFF_MEAN);
  const int r0 = NQ * rq;
  group_bar(rq);
  const int j = tid & 127;
  if (j < NQ) {
    float ss = 0.f, qq = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float2 pr = *reinterpret_cast<const float2*>(part + (((4 * rq + w) * NQ) + j) * 2);
      ss += pr.x;
      qq += pr.y;
    }
    const int ps = pair_slot(r0 + j);
    const float old_mean = means[r0 + j], old_rstd = stats[ps];
    const float sh = use_shift ? old_mean : 0.f;
    const float md = ss * (1.f / 512.f);
    const float var = fmaxf(qq * (1.f / 512.f) - md * md, 0.f);
    const float mu = sh + md, rho = rsqrtf(var + 1e-5f);
    if (CORR) {
      float* gd = reinterpret_cast<float*>(sm + OFF_GD);           // pair layout (row scale, row shift)
      gd[ps] = __fdividef(rho, old_rstd);
      gd[ps + 2] = (old_mean - mu) * rho;
    }
    stats[ps] = rho;
    stats[ps + 2] = -mu * rho;
    means[r0 + j] = mu;
  }
  group_bar(rq);
}

// One channel's 18 rows -> bf16 (hi, lo) operand tile.  Row p = 18*rq + j of a channel lives at
// (pre ^ ((p & 7) << 4)) + 128 * p in the lo image (pre = channel part of the offset, swizzle chunk at bits [4,7)).
// With p & 7 = (2*rq + j) & 7 and 2*rq even, rows j and j^1 differ by bit 4, rows j and j^4 by bit 6, so two
// per-thread bases (y0: j = 0, y2: j = 2) reach every row with one XOR by a constant and an immediate offset.
__device__ __forceinline__ uint32_t row_addr(uint32_t y0, uint32_t y2, int j) {
  const int i = j & 7;
  const uint32_t yb = (i & 2) ? y2 : y0;
  const uint32_t x = (uint32_t)(((i & 1) << 4) | (((i >> 2) & 1) << 6));
  return (x ? (yb ^ x) : yb) + 128u * (uint32_t)((i & 1) + 4 * ((i >> 2) & 1)) + 128u * (uint32_t)(j - i);
}
// Rows are stored in pairs with ONE 32-bit store per image instead of four 16-bit ones: lanes L and L^1 own
// adjacent channels; of a row pair (j, j+4) the even lane stores row j of both channels and the odd lane row j+4
// (one shuffle per pair).  Rows j and j+4 differ in bit 2 of the swizzle, so the two half-warps hit disjoint halves
// of the 32 banks: one wavefront per store.  (Pairing rows j, j+1 would put both halves on the same 16 banks, and
// two lanes writing the 16-bit halves of one word take two wavefronts as well - measured with ncu; the shared-
// memory pipe is this kernel's bottleneck.)  ya0 / ya2: row_addr bases of the EVEN channel of the lane pair for
// rows 0 / 2 of the thread, moved 4 rows down for odd lanes; ytail: address of row 16 (even) / 17 (odd lanes).
template <bool PRECISE>
__device__ __forceinline__ void store_pair(uint32_t addr, bool odd, float u0, float u1) {
  const float recv = __shfl_xor_sync(0xffffffffu, odd ? u0 : u1, 1);
  const float lo_ch = odd ? recv : u0, hi_ch = odd ? u1 : recv;      // channels c (even), c + 1 of this lane's row
  const __nv_bfloat162 hi = __floats2bfloat162_rn(lo_ch, hi_ch);    // .x = lo_ch (low half = lower address)
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hi);
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr + HI_OFF), "r"(hb));
  if (PRECISE) {
    float r0, r1;
    upk2(sub2(pk2(lo_ch, hi_ch), pk2(__uint_as_float(hb << 16), __uint_as_float(hb & 0xFFFF0000u))), r0, r1);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(r0, r1);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&lo)));
  }
}

struct RowBases { uint32_t ya0, ya2, ytail; };

//   MODE 1: (h - mean) * rstd               (provisional normalisation of the channel-mix operand)
//   MODE 2: h                               (output head operand)
template <bool PRECISE, int MODE>
__device__ __forceinline__ void store_rows(const float* hm, const float2* st, uint32_t u_s, const RowBases& rb,
                                           uint32_t ch_off, bool odd, bool ok_tail) {
  // rows j, j+1 normalised as one packed FMA; st + j = (rstd_j, rstd_j+1, -mean*rstd_j, -mean*rstd_j+1) (pair layout)
  auto norm2 = [&](int j, float& n0, float& n1) {
    if (MODE == 2) {
      n0 = hm[j];
      n1 = hm[j + 1];
    } else {
      const float4 s4 = *reinterpret_cast<const float4*>(st + j);
      upk2(fma2(pk2(hm[j], hm[j + 1]), pk2(s4.x, s4.y), pk2(s4.z, s4.w)), n0, n1);
    }
  };
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    if ((j & 4) != 0) continue;              // j = 0, 2, 8, 10: row pairs (j, j+4) and (j+1, j+5)
    float a0, a1, b0, b1;
    norm2(j, a0, a1);
    norm2(j + 4, b0, b1);
    store_pair<PRECISE>(u_s + row_addr(rb.ya0, rb.ya2, j) + ch_off, odd, a0, b0);
    store_pair<PRECISE>(u_s + row_addr(rb.ya0, rb.ya2, j + 1) + ch_off, odd, a1, b1);
  }
  if (ok_tail) {                             // warp-uniform; rows 16, 17
    float a0, a1;
    norm2(16, a0, a1);
    store_pair<PRECISE>(u_s + rb.ytail + ch_off, odd, a0, a1);
  }
}

// LayerNorm 1 of one channel's 18 rows -> the token mix's A operand in tensor memory (this thread's lane, columns
// 9*rq .. 9*rq+8 of the hi and lo images: rows 2j, 2j+1 packed into column j).  The last row quarter also owns the
// tail of the K range: with a spare row (TED) row 70 is the ones row that carries the bias through the GEMM and
// row 71 is zero; rows 72..79 (columns 36..39) are zero - rewritten every time (the head's third accumulator
// buffer overlaps these columns).
template <bool PRECISE, bool ONES_ROW>
__device__ __forceinline__ void store_a_tok(const float* hm, const float2* st, float2 ab, uint32_t taddr_hi, int rq,
                                            bool ok_tail) {
  uint32_t hi[9], lo[9];
#pragma unroll
  for (int j = 0; j < NQ; j += 2) {
    const float4 s4 = *reinterpret_cast<const float4*>(st + j);        // pair layout: (rstd_j, rstd_j+1, nmr_j, nmr_j+1)
    const f32x2 u = fma2(fma2(pk2(hm[j], hm[j + 1]), pk2(s4.x, s4.y), pk2(s4.z, s4.w)), bc2(ab.x), bc2(ab.y));
    float u0, u1;
    upk2(u, u0, u1);
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(u0, u1);           // low half = row j (k even)
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
    hi[j >> 1] = hb;
    if (PRECISE) {
      float l0, l1;
      upk2(sub2(u, pk2(__uint_as_float(hb << 16), __uint_as_float(hb & 0xFFFF0000u))), l0, l1);
      const __nv_bfloat162 l2 = __floats2bfloat162_rn(l0, l1);
      lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
    }
  }
  if (!ok_tail) {                       // rows 16, 17 of this quarter do not exist: (1.0, 0) or zeros
    hi[8] = ONES_ROW ? 0x00003F80u : 0u;
    lo[8] = 0u;
  }
  const uint32_t t_hi = taddr_hi + 9u * (uint32_t)rq;
  tmem_st8(t_hi, hi);
  tmem_st1(t_hi + 8, hi[8]);
  if (PRECISE) {
    tmem_st8(t_hi + TOK_AIMG, lo);
    tmem_st1(t_hi + TOK_AIMG + 8, lo[8]);
  }
  if (rq == 3) {                        // warp-uniform
    tmem_st4(taddr_hi + 36, 0u, 0u, 0u, 0u);
    if (PRECISE) tmem_st4(taddr_hi + TOK_AIMG + 36, 0u, 0u, 0u, 0u);
  }
  tmem_st_wait();
}

// This thread's 18 columns [taddr, taddr + 18) of one accumulator (CAT: plus the columns 72 further that hold
// the other partial product) handed to f(j, value) with compile-time j.  Whole warps only.
template <bool CAT, class F>
__device__ __forceinline__ void acc_rows(uint32_t taddr, F&& f) {     // f(j, packed (v_j, v_j+1)), j even
#pragma unroll
  for (int c0 = 0; c0 < 16; c0 += 8) {
    float a[8], b[8];
    if (CAT) {
      tmem_ld8x2(taddr + c0, taddr + CAT_HI + c0, a, b);
#pragma unroll
      for (int i = 0; i < 8; i += 2) f(c0 + i, add2(pk2(a[i], a[i + 1]), pk2(b[i], b[i + 1])));
    } else {
      tmem_ld8(taddr + c0, a);
#pragma unroll
      for (int i = 0; i < 8; i += 2) f(c0 + i, pk2(a[i], a[i + 1]));
    }
  }
  float a[2], b[2];
  if (CAT) {
    tmem_ld2x2(taddr + 16, taddr + CAT_HI + 16, a, b);
    f(16, add2(pk2(a[0], a[1]), pk2(b[0], b[1])));
  } else {
    tmem_ld2(taddr + 16, a);
    f(16, pk2(a[0], a[1]));
  }
}

// Token-mix accumulator: the 18 columns in ONE load round (16 + 2, a single wait).
template <class F>
__device__ __forceinline__ void acc_rows_tok(uint32_t taddr, F&& f) {  // f(j, packed (v_j, v_j+1)), j even
  float a[16], b[2];
  tmem_ld16p2(taddr, a, b);
#pragma unroll
  for (int i = 0; i < 16; i += 2) f(i, pk2(a[i], a[i + 1]));
  f(16, pk2(b[0], b[1]));
}

// The head reads whole accumulator rows (lane = output feature): columns [0, R) of a channel-type buffer.
template <int R, bool PRECISE, class F>
__device__ __forceinline__ void for_acc_cat(uint32_t taddr, F&& f) {
#pragma unroll
  for (int c0 = 0; c0 < 72; c0 += 8) {
    float a[8], b[8];
    if (PRECISE) {
      tmem_ld8x2(taddr + c0, taddr + CAT_HI + c0, a, b);
    } else {
      tmem_ld8(taddr + CAT_HI + c0, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < R) f(c0 + j, a[j] + b[j]);
  }
}

// Token-mix bias through the GEMM: when the 72-row tile has a spare row (TED: 2S = 70), that row of the
// token-mix operand holds 1.0 (store_a_tok) and column 2S of the block-diagonal weight holds the bias, so the
// epilogue needs no per-row bias load.  BEAT (2S = 72) has no spare row and adds the bias from shared memory.
template <int S>
struct TokBias { static constexpr bool kInGemm = (2 * S + 1 <= 72); };

template <int S, bool PRECISE>
__global__ void LS_CLUSTER_ATTR __launch_bounds__(NT_ALL, 1) fused_step_kernel(const __grid_constant__ FusedParams p) {
  constexpr int R = 2 * S;            // real rows of a tile
  constexpr int NPRE = S - LS_F;      // prefix tokens per pass
  static_assert(R <= 72, "tile rows");
  extern __shared__ uint8_t smem_raw[];
  // 1024-aligned base by POINTER arithmetic on the __shared__ array: an integer round trip would hide the address
  // space from the compiler and turn every shared access below into a generic LD / ST
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
      mbar_init(&bars[BAR_UREADY0 + m], NT_EPI / 32);  // one arrival per epilogue warp (every thread writes a part of M-tile m)
      mbar_init(&bars[BAR_ACC0 + m], 1);
    }
    for (int i = 0; i < 2; ++i) {                 // accumulator buffer i has been read (M-tile i + 2 reuses it)
      mbar_init(&bars[BAR_DRAIN0 + i], NT_EPI / 32);
      mbar_init(&bars[BAR_TDRAIN0 + i], NT_EPI / 32);
    }
    mbar_fence_init();
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  __syncthreads();
  fence_proxy_async_smem();
  tc_fence_before_sync();
#if LS_MULTICAST
  cluster_sync_all();             // barriers of both CTAs initialised before any multicast touches them
#else
  __syncthreads();
#endif
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // Every CTA runs the same number of rounds (with LS_MULTICAST the two CTAs of a cluster consume ONE weight
  // stream and must); a CTA without a real item in the last round recomputes item 0 and writes nothing.
  const int n_items = p.B * p.n_steps;
  const int n_rounds = (n_items + (int)gridDim.x - 1) / (int)gridDim.x;
#if LS_MULTICAST
  const uint32_t cta_rank = cluster_ctarank();
#endif

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
#if LS_NOFETCH
        mbar_arrive_expect_tx_s(full_s, 0);
        continue;
#endif
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
      uint32_t it = 0, uphase = 0, dphase = 0, tphase = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);    // provably warp-uniform copy
      constexpr uint32_t id_cat = idesc_bf16(128, NCAT, 0, 0), id_kk = idesc_bf16(128, NROW, 0, 0);
      // Descriptor words: high word constant per layout, low word = (addr >> 4) | LBO field.
      constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
      const uint32_t u_s = smem_u32(sm + OFF_U);
      const uint32_t uk = desc_lo32(u_s, 16);          // K-major view of block 0 from its lo image
      constexpr uint32_t HI = HI_OFF >> 4;             // descriptor-address distance lo image -> hi image
      constexpr uint32_t BLK = CBS >> 4;               // ... between 64-channel blocks
#if LS_MMA_PROF
      long long prof_ring = 0, prof_u = 0;
      const long long prof_t0 = clock64();
#define LS_PROF(acc, stmt) { const long long t_ = clock64(); stmt; acc += clock64() - t_; }
#else
#define LS_PROF(acc, stmt) { stmt; }
#endif
      auto wait_u_all = [&]() {       // whole operand tile published (K = all channels)
        LS_PROF(prof_u, {
          mbar_wait_s(bars_s + 8 * (BAR_UREADY0 + 0), uphase & 1);
          mbar_wait_s(bars_s + 8 * (BAR_UREADY0 + 1), uphase & 1);
          mbar_wait_s(bars_s + 8 * (BAR_UREADY0 + 2), uphase & 1);
          mbar_wait_s(bars_s + 8 * (BAR_UREADY0 + 3), uphase & 1);
        })
        ++uphase;
        tc_fence_after_sync();
      };
      auto wait_stage = [&]() -> uint32_t {       // -> descriptor low word of the next ring slot
        const uint32_t slot = it % NSLOT, ph = (it / NSLOT) & 1;
        LS_PROF(prof_ring, mbar_wait_s(bars_s + 8 * (BAR_FULL0 + slot), ph);)
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
      // D[buffer] (+)= W[128 x 64 block] * U[:, 64*kc .. +64]^T   (weights = A, K-major; U = B, K-major)
      //   PRECISE: W_hi x [U_lo ; U_hi] -> columns [0,144), then W_lo x U_hi -> columns [72,152)
      //   else   : W_hi x U_hi -> columns [72,152)
      auto gemm_pair = [&](uint32_t d, uint32_t ub, bool first) {
        uint32_t wl = wait_stage();
        if (PRECISE) {
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma_bf16_split_elect(d, wl + 2 * ks, DH, ub + 2 * ks, DH, id_cat, (first && ks == 0) ? 0u : 1u);
          release_stage();
          wl = wait_stage();
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma_bf16_split_elect(d + CAT_HI, wl + 2 * ks, DH, ub + HI + 2 * ks, DH, id_kk, 1u);
          release_stage();
        } else {
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma_bf16_split_elect(d + CAT_HI, wl + 2 * ks, DH, ub + HI + 2 * ks, DH, id_kk, (first && ks == 0) ? 0u : 1u);
          release_stage();
        }
      };
      // one full GEMM over all of U: n_mt M-tiles x n_kc 64-channel blocks.  by_k: the operand tile arrives
      // M-tile by M-tile (channels 128 m .. 128 m + 127 = K blocks 2m, 2m+1), so the first M-tile starts on the
      // K blocks that are already there while the epilogue warps are still storing the rest.
      auto gemm_all = [&](int n_mt, int n_kc, bool by_k, bool head) {
        if (!by_k) wait_u_all();
#pragma unroll 1
        for (int mt = 0; mt < 4; ++mt) {
          if (mt < n_mt) {
            if (!head && mt >= 2) {   // buffer mt - 2 again: its epilogue must have read it
              mbar_wait_s(bars_s + 8 * (BAR_DRAIN0 + mt - 2), dphase & 1);
              tc_fence_after_sync();
            }
            const uint32_t d = tmem_u + (uint32_t)(head ? mt : (mt & 1)) * ACC_COLS;
#pragma unroll 1
            for (int kc = 0; kc < n_kc; ++kc) {
              if (by_k && mt == 0 && (kc & 1) == 0) {
                LS_PROF(prof_u, mbar_wait_s(bars_s + 8 * (BAR_UREADY0 + (kc >> 1)), uphase & 1);)
                tc_fence_after_sync();
              }
              gemm_pair(d, uk + (uint32_t)kc * BLK, kc == 0);
            }
          }
          umma_commit_s_elect(bars_s + 8 * (BAR_ACC0 + mt));    // M-tiles without work still flip their barrier
        }
        if (by_k) ++uphase;
        if (!head) ++dphase;
      };
#pragma unroll 1
      for (int round = 0; round < n_rounds; ++round) {
        gemm_all(4, KIN, false, false);                    // input projection
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
            LS_PROF(prof_u, mbar_wait_s(bars_s + 8 * (BAR_UREADY0 + mt), uphase & 1);)
            if (mt >= 2) mbar_wait_s(bars_s + 8 * (BAR_TDRAIN0 + mt - 2), tphase & 1);   // token accumulator mt - 2 read
            tc_fence_after_sync();
            const uint32_t d = tmem_u + (uint32_t)TOK_D0 + (mt & 1) * NROW;
            const uint32_t a_hi = tmem_u + (uint32_t)TOK_A0 + (mt & 1) * NROW, a_lo = a_hi + TOK_AIMG;
#pragma unroll
            for (uint32_t ks = 0; ks < 5; ++ks) {
              const uint32_t bw = wl[ks >> 2] + 2 * (ks & 3);
              umma_bf16_ts_elect(d, a_hi + 8 * ks, bw, DH, id_kk, ks == 0 ? 0u : 1u);
              if (PRECISE) {
                umma_bf16_ts_elect(d, a_lo + 8 * ks, bw, DH, id_kk, 1u);
                umma_bf16_ts_elect(d, a_hi + 8 * ks, wl[NW - 2 + (ks >> 2)] + 2 * (ks & 3), DH, id_kk, 1u);
              }
            }
            umma_commit_s_elect(bars_s + 8 * (BAR_ACC0 + mt));
          }
          ++uphase;
          ++tphase;
#pragma unroll
          for (int i = 0; i < NW; ++i) release_stage();
          gemm_all(4, 8, true, false);                     // channel mix
        }
        gemm_all(MH, 8, true, true);                       // output head
      }
#if LS_MMA_PROF
      if (p.timing != nullptr && blockIdx.x == 0 && lane == 0) {
        p.timing[508] = prof_ring;
        p.timing[509] = prof_u;
        p.timing[510] = clock64() - prof_t0;
        p.timing[511] = n_rounds;
      }
#endif
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
    // ================= epilogue: thread == (TMEM lane, row quarter) ============================
    const int lq = warp & 3, rq = warp >> 2;
    const int r0 = NQ * rq;                       // first tile row of this thread
    const bool ok_tail = (R == 72) || (rq != 3);  // rows r0+16, r0+17 exist (TED: the last quarter has 16 rows)
    const int c0 = 32 * lq + lane;                // channel of M-tile 0; M-tile m adds 128*m
    const uint32_t lane_base = tmem + ((uint32_t)(lq * 32) << 16);
    float* btok_s = reinterpret_cast<float*>(sm + OFF_BTOK);
    float* emb_s = reinterpret_cast<float*>(sm + OFF_EMB);
    float2* pab_s = reinterpret_cast<float2*>(sm + OFF_PAB);
    float2* psc_s = reinterpret_cast<float2*>(sm + OFF_PSC);
    const float2* stats_q = reinterpret_cast<const float2*>(sm + OFF_STATS) + r0;
    const float2* gd_q = reinterpret_cast<const float2*>(sm + OFF_GD) + r0;
    const uint32_t u_s = smem_u32(sm + OFF_U);
    const bool odd = (lane & 1) != 0;
    RowBases rb;                                  // see row_addr / store_pair
    {
      const int ce = c0 & ~1;                     // even channel of the lane pair
      const uint32_t pre = (uint32_t)(ce >> 6) * CBS + (uint32_t)(((ce & 63) >> 3) << 4) + (uint32_t)(ce & 7) * 2u;
      const uint32_t y0 = (pre ^ ((uint32_t)((2 * rq) & 7) << 4)) + 128u * (uint32_t)r0;
      const uint32_t y2 = (pre ^ ((uint32_t)((2 * rq + 2) & 7) << 4)) + 128u * (uint32_t)(r0 + 2);
      rb.ya0 = odd ? (y0 ^ 64u) + 512u : y0;      // odd lanes store row j + 4 of a pair (j, j+4)
      rb.ya2 = odd ? (y2 ^ 64u) + 512u : y2;
      rb.ytail = (odd ? (y0 ^ 16u) + 128u : y0) + 128u * 16u;     // rows 16 / 17
    }
    uint32_t aphase = 0;
    auto wait_acc = [&](int m) {
      mbar_wait(&bars[BAR_ACC0 + m], aphase & 1);
      __syncwarp();                 // the spin loop may leave the warp diverged; tcgen05.ld is .aligned
      tc_fence_after_sync();
    };
    // One mbarrier arrival per WARP: 32 lanes arriving on the same barrier are 32 serialised shared-memory
    // atomics (ncu: a fifth of all shared wavefronts).  Each lane fences its own writes / TMEM reads, the warp
    // syncs, lane 0 arrives.
    auto publish_u = [&](int m) {   // this warp's part of the operand channels of M-tile m is written
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_UREADY0 + m]);
    };
    auto drained = [&](int bar) {   // this warp's part of an accumulator buffer is read (M-tile m + 2 reuses it)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[bar]);
    };
    float h[72];                    // h[m * 18 + j]: channel c0 + 128 m, row r0 + j
    int tix = 0;

    for (int round = 0; round < n_rounds; ++round) {
      auto stamp = [&]() {
        if (p.timing != nullptr && round == 0 && blockIdx.x == 0 && (tid == 0 || tid == 511) && tix < 256)
          p.timing[(tid ? 256 : 0) + tix++] = clock64();
      };
      const int q_raw = (int)blockIdx.x + round * (int)gridDim.x;
      const bool valid = q_raw < n_items;
      const int q = valid ? q_raw : 0;
      const int k = q / p.B, b = q - k * p.B;           // step within this launch, clip
      const ls_step_params& sp = p.sp[k];
      const StepIO& io = p.io[k];
      const float* x_t = (k == 0) ? p.x_in : p.io[k - 1].x_prev;
      const int t_model = p.t_dev ? (int)min(max(p.t_dev[b], 0ll), (long long)(p.max_t - 1)) : sp.t_model;
      if (k > 0) {                  // x of (k-1, b) comes from another CTA
        if (tid == 0)
          while (ld_acquire_gpu(p.flags + b) < k) __nanosleep(64);
        epi_bar();
      }
      // Per-channel parameters used on the critical path (time embedding, LayerNorm-1 alpha / beta) are parked
      // in shared-memory slots indexed by channel: with 219 KB of shared memory the L1 left over is too small to
      // keep them, and an L2 round trip per M-tile inside the operand stores costs more than the stores.  The four
      // warps that own a channel (rq = 0..3) all store the SAME value into its slot and each thread only reads after
      // its own store, so no barrier is needed; a slot is rewritten for block l + 1 only after the token mix of block
      // l, whose last MMA cannot start before all 16 warps have read the block-l value (racecheck, which does not see
      // the ordering through the MMA, flags these stores; tools/README.md).
      {
        float er[4];
        float2 ab[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          er[m] = p.w.emb_table[(size_t)t_model * LS_D + c0 + 128 * m];
          ab[m] = make_float2(p.w.layer[0].ln1_a[c0 + 128 * m], p.w.layer[0].ln1_b[c0 + 128 * m]);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          emb_s[c0 + 128 * m] = er[m];
          pab_s[c0 + 128 * m] = ab[m];
        }
      }
      // ---- X operand: x_t[b] (hi, lo) into rows p*S + NPRE + f, k = j ----------------------
      const float* xb = x_t + (size_t)b * p.JD * LS_F;
      for (int i = tid; i < p.JD * LS_F; i += NT_EPI) {
        const int j = i / LS_F, f = i - j * LS_F;
        const float v = __ldcg(xb + i);
        store_split<PRECISE>(u_s, tile_off(NPRE + f, j, CBS), v);
        store_split<PRECISE>(u_s, tile_off(S + NPRE + f, j, CBS), v);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) publish_u(m);
      stamp();   // 0: X operand published
      // ---- residual stream init: hoisted terms now, projection result when it lands --------
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int c = c0 + 128 * m;
        const float mu = p.z_mu[(size_t)b * LS_D + c], sd = __expf(0.5f * p.z_lv[(size_t)b * LS_D + c]);
        const float st_c = fmaf(io.eps_c[(size_t)b * LS_D + c], sd, mu), st_u = fmaf(io.eps_u[(size_t)b * LS_D + c], sd, mu);
        const float emo = (NPRE == 2) ? p.emo_tok[(size_t)b * LS_D + c] : 0.f;
        const float* Pb = p.P + (size_t)b * LS_F * LS_D + c;
        const float* Ab = p.A + (size_t)b * LS_F * LS_D + c;
#pragma unroll
        for (int j = 0; j < NQ; ++j) {          // branch-free: the loads of all rows are in flight together
          const int n = r0 + j;
          const bool unc = n >= S;
          const int tok = unc ? n - S : n;
          const int f = min(max(tok - NPRE, 0), LS_F - 1);
          const float pv = Pb[f * LS_D], av = Ab[f * LS_D];
          float v = unc ? pv : pv + av;
          if (NPRE == 2 && tok == 1) v = emo;
          if (tok == 0) v = unc ? st_u : st_c;
          h[m * NQ + j] = (n < R) ? v : 0.f;
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        wait_acc(m);
        acc_rows<PRECISE>(lane_base + (uint32_t)((m & 1) * ACC_COLS + r0) + (PRECISE ? 0u : (uint32_t)CAT_HI),
                          [&](int j, f32x2 v) {
          float v0, v1;
          upk2(v, v0, v1);
          const int n = r0 + j, tok = n >= S ? n - S : n, n1 = n + 1, tok1 = n1 >= S ? n1 - S : n1;
          if (tok >= NPRE && n < R) h[m * NQ + j] += v0;         // prefix-token rows keep their direct values
          if (tok1 >= NPRE && n1 < R) h[m * NQ + j + 1] += v1;
        });
        if (m < 2) drained(BAR_DRAIN0 + m);
      }
      ++aphase;
      stamp();   // 1: input projection consumed
      // ls_debug_hidden (parity aid; off in production: one uniform test per layer).  with_emb: h already carries the
      // next block's time embedding (LS_LN1_EARLY), which is not part of "the residual stream after block l"
      auto dump_hidden = [&](int l, bool with_emb) {
        if (p.dbg_h != nullptr && p.dbg_layer == l && valid && k == 0) {
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
              const int n = r0 + j;
              if (n < R) p.dbg_h[((size_t)b * R + n) * LS_D + c0 + 128 * m] = h[m * NQ + j] - (with_emb ? emb_s[c0 + 128 * m] : 0.f);
            }
        }
      };
      dump_hidden(-1, false);

      for (int l = 0; l < p.n_layers; ++l) {
        const LsLayerW L = p.w.layer[l];
        if (!TokBias<S>::kInGemm && tid < S) btok_s[tid] = btok_s[S + tid] = L.b_tok[tid];
        // x = x + emb ; LN1 ; -> operand tile
        if (!LS_LN1_EARLY || l == 0) {
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const f32x2 emb = bc2(emb_s[c0 + 128 * m]);
#pragma unroll
            for (int j = 0; j < NQ; j += 2) upk2(add2(pk2(h[m * NQ + j], h[m * NQ + j + 1]), emb), h[m * NQ + j], h[m * NQ + j + 1]);
          }
          if (l == 0) ln_stats_q<0>(h, sm, rq, false);     // provisional means for the shift
          ln_stats_q<0>(h, sm, rq, true);
        } else {
          // the previous block's channel-mix epilogue added emb and accumulated the partial sums M-tile by M-tile
          ln_finalize_q<0>(sm, rq, true);
        }
        stamp();   // LN1 stats done
        auto publish_a = [&](int m) {   // this warp's part of the token-mix operand of M-tile m is in tensor memory
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_UREADY0 + m]);
        };
        auto a_tok = [&](int m) {
          store_a_tok<PRECISE, TokBias<S>::kInGemm>(h + m * NQ, stats_q, pab_s[c0 + 128 * m],
                                                    lane_base + (uint32_t)(TOK_A0 + (m & 1) * NROW), rq, ok_tail);
          publish_a(m);
        };
        a_tok(0);
        a_tok(1);
        stamp();   // U1 published (M-tiles 0, 1)
        // token mix epilogue x = x + silu(conv + bias), then straight into the channel-mix operand:
        // normalised with the LN1 statistics; the exact LN2 statistics follow while the GEMM runs
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          wait_acc(m);                  // accumulator m complete => operand buffer m & 1 is free again
          if (m + 2 < 4) a_tok(m + 2);
          acc_rows_tok(lane_base + (uint32_t)(TOK_D0 + (m & 1) * NROW + r0), [&](int j, f32x2 v) {
            if (!TokBias<S>::kInGemm) {
              const float2 b2 = *reinterpret_cast<const float2*>(btok_s + r0 + j);
              v = add2(v, pk2(b2.x, b2.y));
            }
            if (j < 16 || ok_tail) silu_acc2(h[m * NQ + j], h[m * NQ + j + 1], v);
          });
          if (m < 2) drained(BAR_TDRAIN0 + m);
          store_rows<PRECISE, 1>(h + m * NQ, stats_q, u_s, rb, (uint32_t)(2 * m) * CBS, odd, ok_tail);
          publish_u(m);
        }
        ++aphase;
        stamp();   // token-mix epilogue done, U2 published
        ln_stats_q<1>(h, sm, rq, true);
        {     // this layer's channel-mix fold terms and the next layer's LayerNorm-1 parameters: one L2 round trip here,
              // next to the accumulator wait, instead of one per M-tile inside the epilogue (10 % of the stall samples)
          float2 sc[4], ab[4];
          const LsLayerW& Ln = p.w.layer[l + 1 < p.n_layers ? l + 1 : l];
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            sc[m] = make_float2(p.Sc[l * LS_D + c0 + 128 * m], p.tc[l * LS_D + c0 + 128 * m]);
            ab[m] = make_float2(Ln.ln1_a[c0 + 128 * m], Ln.ln1_b[c0 + 128 * m]);
          }
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            psc_s[c0 + 128 * m] = sc[m];
            pab_s[c0 + 128 * m] = ab[m];
          }
        }
        stamp();   // LN2 stats done (under the GEMM)
        // channel mix epilogue: x = x + silu(linear + bias)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          wait_acc(m);
          const float2 sct = psc_s[c0 + 128 * m];
          const float Sc = sct.x, tc = sct.y;
          if (m == 0) stamp();   // first channel-mix accumulator ready
          acc_rows<PRECISE>(lane_base + (uint32_t)((m & 1) * ACC_COLS + r0) + (PRECISE ? 0u : (uint32_t)CAT_HI),
                            [&](int j, f32x2 v) {
            const float4 r = *reinterpret_cast<const float4*>(gd_q + j);     // pair layout: (scale_j, scale_j+1, shift_j, shift_j+1)
            if (j < 16 || ok_tail)
              silu_acc2(h[m * NQ + j], h[m * NQ + j + 1], fma2(pk2(r.x, r.y), v, fma2(pk2(r.z, r.w), bc2(Sc), bc2(tc))));
          });
          if (m < 2) drained(BAR_DRAIN0 + m);
          if (LS_LN1_EARLY && l + 1 < p.n_layers) {
            // this M-tile's channels are final: add the next block's emb and account them in its LayerNorm-1 statistics
            // now, under the MMAs of the later M-tiles, instead of after the last one
            const f32x2 emb = bc2(emb_s[c0 + 128 * m]);
#pragma unroll
            for (int j = 0; j < NQ; j += 2) upk2(add2(pk2(h[m * NQ + j], h[m * NQ + j + 1]), emb), h[m * NQ + j], h[m * NQ + j + 1]);
            if (m == 0)
              ln_partial_q<true>(h + m * NQ, sm, rq);
            else
              ln_partial_q<false>(h + m * NQ, sm, rq);
          }
          if (m == 2) stamp();   // three of four M-tiles consumed
        }
        ++aphase;
        stamp();   // channel-mix epilogue done
        dump_hidden(l, LS_LN1_EARLY && l + 1 < p.n_layers);
      }

      // ---- output head operand: plain hi/lo split of h -------------------------------------
      // This thread has waited for every M-tile of the last channel mix, so all MMAs reading U are complete.
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        store_rows<PRECISE, 2>(h + m * NQ, stats_q, u_s, rb, (uint32_t)(2 * m) * CBS, odd, ok_tail);
        publish_u(m);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) wait_acc(m);
      ++aphase;
      // head epilogue with the lane = output feature view: warp group rq reads M-tile rq, all rows
      if ((rq * 4 + lq) * 32 < p.JD) {           // warp-uniform: tcgen05.ld is .aligned
        // h is dead: reuse it for the head outputs (lane = output feature, column = row)
        for_acc_cat<R, PRECISE>(lane_base + (uint32_t)((rq % NACC) * ACC_COLS), [&](int n, float v) { h[n] = v; });
        const int c = 128 * rq + c0;              // output feature
        if (c < p.JD && valid) {
          const float bo = p.w.b_out[c], sc = p.scale[b];
          const size_t base = ((size_t)b * p.JD + c) * LS_F;
          const float* nzp = io.noise ? io.noise + (size_t)b * io.nsb + (size_t)c * io.nsj : nullptr;
          const bool use_noise = (sp.mode != 2) && sp.add_noise && nzp != nullptr;
          // guidance first (frees the uncond half), then the sampler update in two batches whose global LOADS are
          // all issued before the first STORE: x_prev may alias x_t, so the compiler cannot hoist a load above a
          // store and a load-per-frame loop would pay one L2 round trip per frame on a single warp.
#pragma unroll
          for (int f = 0; f < LS_F; ++f) {
            const float oc = h[NPRE + f] + bo, ou = h[S + NPRE + f] + bo;
            h[NPRE + f] = ou + sc * (oc - ou);                    // cfg_sampler.py:31
          }
#pragma unroll
          for (int f0 = 0; f0 < LS_F; f0 += 17) {
            float xt[17], nz[17];
#pragma unroll
            for (int i = 0; i < 17; ++i) {
              xt[i] = __ldcg(x_t + base + f0 + i);
              nz[i] = use_noise ? nzp[(size_t)(f0 + i) * io.nsf] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 17; ++i) {
              float xp = 0.f;
              const float x0 = ls_sampler_update(sp, h[NPRE + f0 + i], xt[i], nz[i], &xp);
              if (io.pred_x0) io.pred_x0[base + f0 + i] = x0;
              if (sp.mode != 2) io.x_prev[base + f0 + i] = xp;
            }
          }
        }
      }
      // the next item's X operand overwrites U and its accumulators the head's: all reads above are done
      if (p.n_steps > 1) __threadfence();        // x_prev visible device-wide before the counter moves
      tc_fence_before_sync();
      epi_bar();
      if (p.n_steps > 1 && valid && tid == 0) st_release_gpu(p.flags + b, k + 1);
    }
  }
  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before_sync();
#if LS_MULTICAST
  cluster_sync_all();             // the peer may still multicast commits into this CTA's barriers
#else
  __syncthreads();
#endif
  if (warp == 17) tmem_dealloc<512>(tmem);
}

// ---- weight tape ------------------------------------------------------------------------------
// stage = [128 x 64] block of W (rows m0.., cols k0..) as K-major swizzle-128B images, hi then lo.
// kscale (or nullptr) multiplies column k: the channel-mix weights carry LayerNorm 2's alpha.
__global__ void build_w_stage_kernel(const float* __restrict__ src, int rows, int cols, int ld, int m0, int k0,
                                     const float* __restrict__ kscale, uint8_t* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 128 * 64; i += gridDim.x * blockDim.x) {
    const int m = i >> 6, k = i & 63;
    float v = (m0 + m < rows && k0 + k < cols) ? src[(size_t)(m0 + m) * ld + k0 + k] : 0.f;
    if (kscale != nullptr && k0 + k < cols) v = __fmul_rn(v, kscale[k0 + k]);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const uint32_t off = tile_off(m, k, 0);
    *reinterpret_cast<__nv_bfloat16*>(dst + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(dst + W_HALF + off) = lo;
  }
}

// Per output channel c of a channel-mix layer (one warp each):
//   Sc[c] = sum_k fl(W[c,k]*alpha[k]),  tc[c] = sum_k W[c,k]*beta[k] + bias[c]   (fp64 accumulation)
__global__ void ch_fold_kernel(const float* __restrict__ w, const float* __restrict__ alpha, const float* __restrict__ beta,
                               const float* __restrict__ bias, float* __restrict__ Sc, float* __restrict__ tc) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= LS_D) return;
  double s = 0.0, t = 0.0;
  for (int k = lane; k < LS_D; k += 32) {
    const float wv = w[(size_t)c * LS_D + k];
    s += (double)__fmul_rn(wv, alpha[k]);
    t += (double)wv * (double)beta[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  if (lane == 0) {
    Sc[c] = (float)s;
    tc[c] = (float)(t + (double)bias[c]);
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
  float* Sc = nullptr;     // [n_layers][512]
  float* tc = nullptr;
  int* flags = nullptr;    // [max_batch]
  int KIN = 0, MH = 0, n_stages = 0;
  int sm_count = 0;
  bool attr_done[2][2] = {{false, false}, {false, false}};
  int max_coresident[2][2] = {{0, 0}, {0, 0}};
  long long* tbuf = nullptr;   // LS_FUSED_TIMING stamps
};

template <int S, bool PRECISE>
int launch_fused(ls_handle* h, FusedState* fs, const FusedParams& fp, cudaStream_t s) {
  bool& done = fs->attr_done[S - 35][PRECISE ? 1 : 0];
  if (!done) {
    LS_CUDA(h, cudaFuncSetAttribute(fused_step_kernel<S, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DYN));
    done = true;
  }
  // at most one CTA per SM; cluster pairs need an even grid.  max_coresident = occupancy x SMs of THIS context
  // (MPS / green-context SM limits included): the grid of a multi-step launch never exceeds it.
  const int items = fp.B * fp.n_steps;
  if (fs->max_coresident[S - 35][PRECISE ? 1 : 0] == 0) {
    int per_sm = 0;
    LS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_step_kernel<S, PRECISE>, NT_ALL, SMEM_DYN));
    if (per_sm < 1) return ls_fail(h, LS_EUNSUPPORTED, "the fused kernel does not fit on an SM of this device");
    fs->max_coresident[S - 35][PRECISE ? 1 : 0] = per_sm * fs->sm_count;
  }
#if LS_MULTICAST
  const int grid = (items + 1 < fs->sm_count ? items + 1 : fs->sm_count) & ~1;
#else
  const int grid = items < fs->sm_count ? items : fs->sm_count;
#endif
  if (fp.n_steps > 1 && grid > fs->max_coresident[S - 35][PRECISE ? 1 : 0])
    return ls_fail(h, LS_EUNSUPPORTED, "multi-step launch needs %d co-resident CTAs, the device offers %d", grid,
                   fs->max_coresident[S - 35][PRECISE ? 1 : 0]);
  if (fp.n_steps > 1) {
    // CTAs of a multi-step launch wait for each other (per-clip step counters), so the whole grid must be
    // co-resident: a cooperative launch guarantees it (or fails with cudaErrorCooperativeLaunchTooLarge) even when
    // other kernels hold SMs - a plain launch could leave producers unscheduled behind spinning consumers.
    void* args[] = {const_cast<FusedParams*>(&fp)};
    LS_CUDA(h, cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&fused_step_kernel<S, PRECISE>), dim3(grid),
                                           dim3(NT_ALL), args, SMEM_DYN, s));
    h->launches++;
    return LS_OK;
  }
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
    if (fs->Sc) cudaFree(fs->Sc);
    if (fs->tc) cudaFree(fs->tc);
    if (fs->flags) cudaFree(fs->flags);
    if (fs->tbuf) cudaFree(fs->tbuf);
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
    if (fs->KIN > 8 || fs->MH > NACC) {      // the head's M-tiles each need their own accumulator buffer
      delete fs;
      return 1;
    }
    fs->n_stages = 8 * fs->KIN + h->cfg.n_layers * 68 + 16 * fs->MH;
    fs->tape_bytes = (size_t)fs->n_stages * SLOT;
    const size_t fold_bytes = (size_t)h->cfg.n_layers * LS_D * sizeof(float);
    if (cudaMalloc(&fs->tape, fs->tape_bytes) != cudaSuccess || cudaMalloc(&fs->Sc, fold_bytes) != cudaSuccess ||
        cudaMalloc(&fs->tc, fold_bytes) != cudaSuccess ||
        cudaMalloc(&fs->flags, (size_t)h->cfg.max_batch * sizeof(int)) != cudaSuccess) {
      if (fs->tape) cudaFree(fs->tape);
      if (fs->Sc) cudaFree(fs->Sc);
      if (fs->tc) cudaFree(fs->tc);
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
      build_w_stage_kernel<<<16, 256, 0, s>>>(win, LS_D, h->JD, IN, mt * 128, kc * 64, nullptr, fs->tape + st * SLOT);
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
    const LsLayerW& L = h->w.layer[l];
    ch_fold_kernel<<<LS_D / 8, 256, 0, s>>>(wch, L.ln2_a, L.ln2_b, L.b_ch, fs->Sc + (size_t)l * LS_D, fs->tc + (size_t)l * LS_D);
    LS_LAUNCH_CHECK(h);
    for (int mt = 0; mt < 4; ++mt)
      for (int kc = 0; kc < 8; ++kc) {
        build_w_stage_kernel<<<16, 256, 0, s>>>(wch, LS_D, LS_D, LS_D, mt * 128, kc * 64, L.ln2_a, fs->tape + st * SLOT);
        LS_LAUNCH_CHECK(h);
        st += 2;
      }
  }
  const float* wout = raw("output_process.poseFinal.weight");
  for (int mt = 0; mt < fs->MH; ++mt)
    for (int kc = 0; kc < 8; ++kc) {
      build_w_stage_kernel<<<16, 256, 0, s>>>(wout, h->JD, LS_D, LS_D, mt * 128, kc * 64, nullptr, fs->tape + st * SLOT);
      LS_LAUNCH_CHECK(h);
      st += 2;
    }
  if ((int)st != fs->n_stages) return ls_fail(h, LS_EINVAL, "tape stage count mismatch");
  return LS_OK;
}

// n_steps consecutive steps in one launch (n_steps <= LS_MAX_FUSED_STEPS).  Step k reads x from
// x_in (k = 0) or io[k-1].x_prev, so with n_steps > 1 every step needs its own x_prev buffer.
int lsf_steps(ls_handle* h, int B, int n_steps, const ls_step_params* p, const ls_step_io* io, int precise,
              const float* x_in, const float* scale, cudaStream_t s, const int64_t* t_dev) {
  FusedState* fs = static_cast<FusedState*>(h->fused);
  if (!fs) return ls_fail(h, LS_EUNSUPPORTED, "tcgen05 path not available");
  if (n_steps < 1 || n_steps > KMAX) return ls_fail(h, LS_EINVAL, "n_steps %d outside [1,%d]", n_steps, KMAX);
  if (LS_MULTICAST && n_steps > 1 && B < 2) {
    // With one clip the two CTAs of a cluster would hold consecutive steps of the SAME clip: the second waits
    // for the first, which in turn needs its peer to drain the shared weight ring.  Run step by step.
    for (int k = 0; k < n_steps; ++k) {
      const int rc = lsf_steps(h, B, 1, p + k, io + k, precise, k == 0 ? x_in : io[k - 1].x_prev, scale, s, t_dev);
      if (rc != LS_OK) return rc;
    }
    return LS_OK;
  }
  FusedParams fp{};
  fp.tape = fs->tape;
  fp.n_layers = h->cfg.n_layers;
  fp.JD = h->JD;
  fp.KIN = fs->KIN;
  fp.MH = fs->MH;
  fp.B = B;
  fp.n_steps = n_steps;
  fp.w = h->w;
  fp.Sc = fs->Sc; fp.tc = fs->tc;
  fp.A = h->A; fp.P = h->P; fp.z_mu = h->z_mu; fp.z_lv = h->z_lv; fp.emo_tok = h->emo_tok;
  fp.x_in = x_in; fp.scale = scale;
  fp.flags = fs->flags;
  fp.t_dev = reinterpret_cast<const long long*>(t_dev);
  fp.max_t = h->cfg.max_timestep;
  fp.dbg_h = h->dbg_h;
  fp.dbg_layer = h->dbg_layer;
  for (int k = 0; k < n_steps; ++k) {
    fp.sp[k] = p[k];
    fp.io[k].eps_c = io[k].eps_cond; fp.io[k].eps_u = io[k].eps_uncond; fp.io[k].noise = io[k].noise;
    fp.io[k].nsb = io[k].noise_sb; fp.io[k].nsj = io[k].noise_sj; fp.io[k].nsf = io[k].noise_sf;
    fp.io[k].x_prev = io[k].x_prev; fp.io[k].pred_x0 = io[k].pred_x0;
  }
  if (n_steps > 1) LS_CUDA(h, cudaMemsetAsync(fs->flags, 0, (size_t)B * sizeof(int), s));
  fp.timing = nullptr;
  long long*& tbuf = fs->tbuf;        // per handle (= per device)
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
    static const char* names[] = {"LN1 stats", "U1 publish", "tok epilogue + U2 publish", "LN2 stats", "ch acc 0 wait",
                                  "ch epilogue m0-2", "ch epilogue m3"};
#if LS_MMA_PROF
    fprintf(stderr, "[fused timing] MMA warp of CTA 0 over %lld rounds: total %lld cyc, waiting for weight stages %lld, "
                    "waiting for operand tiles %lld\n", t[511], t[510], t[508], t[509]);
#endif
    for (int w = 0; w < 2; ++w) {
      const long long* q = t + 256 * w;
      fprintf(stderr, "[fused timing] thread %d: X publish -> in-proj consumed %lld cyc\n", w ? 511 : 0, q[1] - q[0]);
      for (int l = 0; l < h->cfg.n_layers && 2 + 7 * l + 6 < 256; ++l) {
        fprintf(stderr, "  layer %d:", l);
        for (int k = 0; k < 7; ++k) fprintf(stderr, " %s %lld |", names[k], q[2 + 7 * l + k] - q[1 + 7 * l + k]);
        fprintf(stderr, "\n");
      }
    }
  }
  return rc;
}
