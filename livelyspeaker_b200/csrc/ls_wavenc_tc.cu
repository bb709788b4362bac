// WavEncoder layers 2-4 (audio_enc.py:12-18: Conv1d(k=15, stride 6), 32->64->128->256 channels) as implicit GEMMs
// on the tcgen05 tensor cores, bf16x3 split operands with fp32 accumulation.
//
//   out[b, co, lo] = bias[co] + sum_{ci, k<15} w[co, ci, k] * in[b, ci, lo*6 + k]
//   D[position, co] = A[position, (ci, k)] * B[co, (ci, k)]^T        M = 128 positions, N = Co, K = Ci*16
//
// K packs 16 taps per input channel (tap 15 has a zero weight), so one im2col row of one channel is 16
// CONSECUTIVE input samples = two 16-byte chunks of the swizzle-128B K-major tile: every thread builds the row
// of its own output position straight from global memory (hi / lo bf16 images) with 128-bit shared stores.
// The weights come pre-swizzled from a tape (one bulk async copy per K chunk of 4 input channels).  Two stages:
// the im2col build of chunk c+1 overlaps the MMAs of chunk c.  One CTA = 128 output positions of one clip.
// Layer 1 (1 input channel, 4 % of the flops) and InstanceNorm + LeakyReLU stay on the CUDA-core kernels.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>

#include "ls_internal.cuh"
#include "ls_rows_gemm.cuh"
#include "ls_tc.cuh"

using namespace lstc;

namespace {

constexpr int CONV_K = 15, KC = 64, CI_PER_CHUNK = 4, STRIDE = 6;
constexpr uint32_t A_IMG = 128 * 128;            // 128 positions x 64 bf16

template <int CO>
struct Lay {
  static constexpr uint32_t B_IMG = CO * 128;     // CO rows x 64 bf16
  static constexpr uint32_t STAGE = 2 * A_IMG + 2 * B_IMG;
  static constexpr uint32_t SMEM = 2 * STAGE + 1024 + 64;
};

__device__ __forceinline__ uint32_t pack_hi_lo(float a, float b, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hb << 16), b - __uint_as_float(hb & 0xFFFF0000u));
  *lo = *reinterpret_cast<const uint32_t*>(&l);
  return hb;
}

template <int CO>
__global__ void __launch_bounds__(128, 1) wav_conv_tc_kernel(const float* __restrict__ in, const uint8_t* __restrict__ tape,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             int Ci, int Li, int Lo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * Lay<CO>::STAGE);     // full[2], empty[2], acc
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + 2 * Lay<CO>::STAGE + 48);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y, lo = blockIdx.x * 128 + tid;
  const bool valid = lo < Lo;
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<CO>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sm_s = smem_u32(sm), bars_s = smem_u32(bars);
  const int n_chunks = Ci / CI_PER_CHUNK;
  const float* src0 = in + (size_t)b * Ci * Li + (size_t)lo * STRIDE;
  const uint32_t row_off = (uint32_t)(tid >> 3) * 1024u + (uint32_t)(tid & 7) * 128u;
  constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
  constexpr uint32_t idesc = idesc_bf16(128, CO, 0, 0);

  for (int c = 0; c < n_chunks; ++c) {
    const int s = c & 1;
    uint8_t* stage = sm + s * Lay<CO>::STAGE;
    if (c >= 2) {                                     // the MMAs that read this stage (chunk c-2) are done
      mbar_wait(&bars[2 + s], ((c >> 1) - 1) & 1);
      tc_fence_after_sync();
    }
    if (tid == 0) {                                   // weights of this chunk: hi image then lo image, contiguous
      mbar_arrive_expect_tx(&bars[s], 2 * Lay<CO>::B_IMG);
      bulk_g2s(stage + 2 * A_IMG, tape + (size_t)c * 2 * Lay<CO>::B_IMG, 2 * Lay<CO>::B_IMG, &bars[s]);
    }
    // im2col row of this thread's position: 4 input channels x 16 samples
#pragma unroll
    for (int cil = 0; cil < CI_PER_CHUNK; ++cil) {
      const float* src = src0 + (size_t)(c * CI_PER_CHUNK + cil) * Li;
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = (valid && lo * STRIDE + k < Li) ? __ldg(src + k) : 0.f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint4 hi, lw;
        hi.x = pack_hi_lo(v[8 * hh + 0], v[8 * hh + 1], &lw.x);
        hi.y = pack_hi_lo(v[8 * hh + 2], v[8 * hh + 3], &lw.y);
        hi.z = pack_hi_lo(v[8 * hh + 4], v[8 * hh + 5], &lw.z);
        hi.w = pack_hi_lo(v[8 * hh + 6], v[8 * hh + 7], &lw.w);
        const uint32_t off = row_off + ((uint32_t)((cil * 2 + hh) ^ (tid & 7)) << 4);
        *reinterpret_cast<uint4*>(stage + off) = hi;
        *reinterpret_cast<uint4*>(stage + A_IMG + off) = lw;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {                                  // whole warp in uniform control flow, one elected lane issues
      mbar_wait_s(bars_s + 8 * s, (c >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t a_hi = desc_lo32(sm_s + s * Lay<CO>::STAGE, 16), a_lo = desc_lo32(sm_s + s * Lay<CO>::STAGE + A_IMG, 16);
      const uint32_t b_hi = desc_lo32(sm_s + s * Lay<CO>::STAGE + 2 * A_IMG, 16),
                     b_lo = desc_lo32(sm_s + s * Lay<CO>::STAGE + 2 * A_IMG + Lay<CO>::B_IMG, 16);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
#pragma unroll
      for (uint32_t ks = 0; ks < 4; ++ks) {
        umma_bf16_split_elect(tm, a_hi + 2 * ks, DH, b_hi + 2 * ks, DH, idesc, (c > 0 || ks > 0) ? 1u : 0u);
        umma_bf16_split_elect(tm, a_lo + 2 * ks, DH, b_hi + 2 * ks, DH, idesc, 1u);
        umma_bf16_split_elect(tm, a_hi + 2 * ks, DH, b_lo + 2 * ks, DH, idesc, 1u);
      }
      umma_commit_s_elect(bars_s + 8 * (2 + s));
      if (c == n_chunks - 1) umma_commit_s_elect(bars_s + 8 * 4);
    }
  }
  mbar_wait(&bars[4], 0);
  __syncwarp();
  tc_fence_after_sync();
  // epilogue: lane = output position, column = output channel
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  float* dst = out + (size_t)b * CO * Lo + lo;
#pragma unroll 1
  for (int c0 = 0; c0 < CO; c0 += 16) {
    float v[16];
    tmem_ld16(taddr + c0, v);
    if (valid) {
#pragma unroll
      for (int j = 0; j < 16; ++j) dst[(size_t)(c0 + j) * Lo] = v[j] + __ldg(bias + c0 + j);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<CO>(tmem);
}

// tape of one layer: per K chunk (4 input channels) the [Co x 64] weight tile as K-major swizzle-128B images, hi then lo
__global__ void build_conv_tape_kernel(const float* __restrict__ w, int Co, int Ci, uint8_t* __restrict__ dst) {
  const int n = Co * Ci * 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int k = i & 15, ci = (i >> 4) % Ci, co = i / (16 * Ci);
    const float v = k < CONV_K ? w[((size_t)co * Ci + ci) * CONV_K + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const int c = ci / CI_PER_CHUNK, kk = (ci % CI_PER_CHUNK) * 16 + k;
    uint8_t* base = dst + (size_t)c * 2 * Co * 128;
    const uint32_t off = tile_off(co, kk, 0);
    *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + (size_t)Co * 128 + off) = lo;
  }
}

struct WavTc {
  uint8_t* tape[3] = {nullptr, nullptr, nullptr};
  uint8_t* tape_a = nullptr;      // audio half of input_mapping (W_a [512 x 256]) for lsw_audio_proj
  bool attr_done = false;
  // lsw_encoder_fused: per-tile partial sums and (mean, rstd) rows, grown on demand
  void* v2_part = nullptr;
  void* v2_stats = nullptr;
  float w1_host[32 * 15] = {};     // layer-1 weights as a kernel parameter
  size_t v2_part_elems = 0, v2_stats_elems = 0;
  bool v2_attr_done = false;
};

template <int CO>
int launch_conv(ls_handle* h, const float* in, const uint8_t* tape, const float* bias, float* out, int nb, int Ci, int Li,
                int Lo, cudaStream_t s) {
  wav_conv_tc_kernel<CO><<<dim3((Lo + 127) / 128, nb), 128, Lay<CO>::SMEM, s>>>(in, tape, bias, out, Ci, Li, Lo);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

}  // namespace

int lsw_available(const ls_handle* h) { return h && h->wavtc != nullptr; }

void lsw_destroy(ls_handle* h) {
  if (h && h->wavtc) {
    WavTc* w = static_cast<WavTc*>(h->wavtc);
    for (auto& t : w->tape)
      if (t) cudaFree(t);
    if (w->tape_a) cudaFree(w->tape_a);
    if (w->v2_part) cudaFree(w->v2_part);
    if (w->v2_stats) cudaFree(w->v2_stats);
    delete w;
    h->wavtc = nullptr;
  }
}

// w[i]: conv weights of layers 2..4 ([Co, Ci, 15]); builds the three tapes
int lsw_init(ls_handle* h, const float* const w[3], cudaStream_t s, const float* w0) {
  static const int CO[3] = {64, 128, 256}, CI[3] = {32, 64, 128};
  WavTc* wt = static_cast<WavTc*>(h->wavtc);
  if (!wt) {
    wt = new WavTc();
    for (int i = 0; i < 3; ++i)
      if (cudaMalloc(&wt->tape[i], (size_t)CO[i] * CI[i] * 16 * 4) != cudaSuccess) {
        for (auto& t : wt->tape)
          if (t) cudaFree(t);
        delete wt;
        return ls_fail(h, LS_ENOMEM, "WavEncoder weight tape");
      }
    if (cudaMalloc(&wt->tape_a, lsrg::rows_tape_bytes(LS_D, LS_AF)) != cudaSuccess) {
      for (auto& t : wt->tape) cudaFree(t);
      delete wt;
      return ls_fail(h, LS_ENOMEM, "audio projection weight tape");
    }
    h->wavtc = wt;
  }
  if (!wt->attr_done) {
    LS_CUDA(h, cudaFuncSetAttribute(wav_conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay<64>::SMEM));
    LS_CUDA(h, cudaFuncSetAttribute(wav_conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay<128>::SMEM));
    LS_CUDA(h, cudaFuncSetAttribute(wav_conv_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay<256>::SMEM));
    LS_CUDA(h, cudaFuncSetAttribute(lsrg::rows_gemm_kernel<lsrg::AChanMajor34, lsrg::EpiStore<false>>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsrg::SMEM));
    wt->attr_done = true;
  }
  lsrg::build_rows_tape_kernel<<<64, 256, 0, s>>>(h->w.w_a_t, LS_D, LS_AF, wt->tape_a);
  LS_LAUNCH_CHECK(h);
  if (w0 != nullptr) {           // layer-1 weights [32][1][15] to the host once: they are passed as a kernel parameter
    LS_CUDA(h, cudaMemcpyAsync(wt->w1_host, w0, sizeof(wt->w1_host), cudaMemcpyDeviceToHost, s));
    LS_CUDA(h, cudaStreamSynchronize(s));
  }
  for (int i = 0; i < 3; ++i) {
    build_conv_tape_kernel<<<64, 256, 0, s>>>(w[i], CO[i], CI[i], wt->tape[i]);
    LS_LAUNCH_CHECK(h);
  }
  return LS_OK;
}

// layer = 0, 1, 2 for the 2nd, 3rd, 4th convolution; in [nb, Ci, Li] -> out [nb, Co, Lo] (bias added, no activation)
int lsw_conv(ls_handle* h, int layer, const float* in, const float* bias, float* out, int nb, int Li, int Lo, cudaStream_t s) {
  WavTc* wt = static_cast<WavTc*>(h->wavtc);
  if (!wt) return ls_fail(h, LS_EUNSUPPORTED, "tensor-core WavEncoder not initialised");
  switch (layer) {
    case 0: return launch_conv<64>(h, in, wt->tape[0], bias, out, nb, 32, Li, Lo, s);
    case 1: return launch_conv<128>(h, in, wt->tape[1], bias, out, nb, 64, Li, Lo, s);
    case 2: return launch_conv<256>(h, in, wt->tape[2], bias, out, nb, 128, Li, Lo, s);
  }
  return ls_fail(h, LS_EINVAL, "layer %d", layer);
}

// A[b, f, :] = af[b, :, f] . W_a^T (RAG.py:184-192, the audio columns of input_mapping; cond pass only): one GEMM over
// the nb * 34 frame rows of a WavEncoder chunk, A operand read in the encoder's own channel-major layout.
int lsw_audio_proj(ls_handle* h, const float* af_cm, int nb, float* A, cudaStream_t s) {
  WavTc* wt = static_cast<WavTc*>(h->wavtc);
  if (!wt) return ls_fail(h, LS_EUNSUPPORTED, "tensor-core precompute not initialised");
  const int rows = nb * LS_F;
  lsrg::rows_gemm_kernel<<<dim3(lsrg::row_tiles(rows), 1), lsrg::NTHREADS, lsrg::SMEM, s>>>(
      lsrg::AChanMajor34{af_cm, LS_AF}, wt->tape_a, rows, LS_AF, lsrg::EpiStore<false>{A, LS_D, nullptr});
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

// =====================================================================================================================
// Round-2 WavEncoder pipeline: InstanceNorm fused away.
//   audio_enc.py:9-19: Conv1d -> InstanceNorm1d -> LeakyReLU(0.3) three times, then a last Conv1d.  InstanceNorm needs the
//   statistics of a whole (clip, channel) row, so the round-1 pipeline ran it as a separate pass over every layer's
//   output (405 us of HBM traffic at B = 512).  Here
//     * every conv writes its RAW accumulators (no bias: InstanceNorm removes a per-channel constant anyway) and, per
//       128-position tile, the partial sums (sum a, sum a^2) of each output channel (deterministic: transposing warp
//       butterflies + a fixed-order combine),
//     * in_finalize_kernel turns the partials of a row into (mean, rstd) in fp64,
//     * the NEXT conv applies (a - mean) * rstd and the LeakyReLU while it loads its input window.
//   The tensor-core convs also load their input differently: the window of a 128-position tile is contiguous in memory
//   (778 floats per input channel), so the CTA reads it coalesced into a shared fp32 staging row - normalising each
//   element once - and the threads build their overlapping 16-tap im2col rows from shared memory.  The round-1 loader
//   gathered 64 scalars per thread per chunk at a 24-byte stride (L1-wavefront-bound: 450 per warp per chunk).
// =====================================================================================================================
namespace {

constexpr int C1 = 32, C1_TILE = 256, C1_POS = 4, C1_SPAN = C1_TILE * C1_POS, C1_STRIDE = 5, C1_PAD = 1600;

// V values per lane in, lane idx holds the sum over the 32 lanes of value idx (V = 32: idx = lane)
__device__ __forceinline__ void butterfly32(float* v, int lane) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int xm = 16 >> s, half = 16 >> s;
    const bool up = (lane & xm) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? v[j] : v[j + half];
      const float keep = up ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, xm);
    }
  }
}

// layer 1: audio [B][L0] -> raw accumulators [B][32][L1] + partials part[b][c][tile] = (sum, sum of squares).
// The 480 weights travel as a kernel PARAMETER (constant bank): every FMA takes its weight as a constant operand.  With
// the weights in shared memory the kernel issued one LDS per FMA and was LSU-bound at 266 us for 517 MB of output.
// Instruction-issue-bound (75 % issue-active at 223 us with one position per thread; see the loop below).
struct Conv1W { float w[C1 * CONV_K]; };
__global__ void __launch_bounds__(C1_TILE, 2) wav_conv1_kernel(const float* __restrict__ audio, const __grid_constant__ Conv1W cw,
                                                            float* __restrict__ out, float2* __restrict__ part, int L0, int L1,
                                                            int n_tiles) {
  constexpr int XW = (C1_SPAN - 1) * C1_STRIDE + CONV_K;
  __shared__ float xs[XW + 1];
  __shared__ float wpart[C1_TILE / 32][2 * C1];
  const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lo0 = tile * C1_SPAN;
  const float* ab = audio + (size_t)b * L0;
  const int x0 = lo0 * C1_STRIDE - C1_PAD;
  for (int i = tid; i < XW; i += C1_TILE) {
    const int g = x0 + i;
    xs[i] = (g >= 0 && g < L0) ? ab[g] : 0.f;
  }
  __syncthreads();
  // A thread takes C1_POS positions (lo0 + tid + 256 p), one after the other, and keeps the running sum and sum of
  // squares of its 32 channels: the two 32-value butterflies below are paid once per thread, not once per position
  // (they were a fifth of the instruction stream): 224 -> 204 us.  Sequential on purpose (unroll 1) and capped at 128
  // registers: unrolled, the compiler keeps every position's accumulators live; uncapped, it hoists the 480 weights
  // out of the loop into 254 registers.  Two passes of 16 channels (one butterfly each, 64-80 registers) measured slower.
  float acc[C1], sq[C1];
#pragma unroll
  for (int c = 0; c < C1; ++c) acc[c] = sq[c] = 0.f;
#pragma unroll 1
  for (int p = 0; p < C1_POS; ++p) {
    const int lp = tid + C1_TILE * p, lo = lo0 + lp;
    if (lo >= L1) break;
    float x[CONV_K];
#pragma unroll
    for (int k = 0; k < CONV_K; ++k) x[k] = xs[lp * C1_STRIDE + k];
    float* dst = out + (size_t)b * C1 * L1 + lo;
#pragma unroll
    for (int c = 0; c < C1; ++c) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < CONV_K; ++k) a = fmaf(cw.w[c * CONV_K + k], x[k], a);
      dst[(size_t)c * L1] = a;
      acc[c] += a;
      sq[c] = fmaf(a, a, sq[c]);
    }
  }
  butterfly32(acc, lane);
  butterfly32(sq, lane);
  wpart[warp][lane] = acc[0];
  wpart[warp][C1 + lane] = sq[0];
  __syncthreads();
  if (tid < 2 * C1) {
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < C1_TILE / 32; ++wq) t += wpart[wq][tid];
    float* dst = reinterpret_cast<float*>(part + ((size_t)b * C1 + (tid & (C1 - 1))) * n_tiles + tile);
    dst[tid >> 5] = t;
  }
}

// (mean, rstd) of every (clip, channel) row from its per-tile partial sums; fp64, fixed order
__global__ void in_finalize_kernel(const float2* __restrict__ part, int n_rows, int n_tiles, int L, float2* __restrict__ stats) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  double s = 0.0, q = 0.0;
  for (int t = 0; t < n_tiles; ++t) {
    const float2 p = part[(size_t)r * n_tiles + t];
    s += p.x;
    q += p.y;
  }
  const double mean = s / L, var = fmax(q / L - mean * mean, 0.0);
  stats[r] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
}

template <int CO>
struct Lay2 {
  static constexpr uint32_t B_IMG = CO * 128;
  static constexpr uint32_t STAGE = 2 * A_IMG + 2 * B_IMG;
  static constexpr uint32_t RAW_ROW = 784;                             // floats per staged input channel (778 used)
  static constexpr uint32_t OFF_RAW = 2 * STAGE;
  static constexpr uint32_t OFF_WPART = OFF_RAW + 8 * CI_PER_CHUNK * 108 * 4;     // 8 warps x 4 channels x 108 floats; then [4 warps][2 * CO] epilogue partials
  static constexpr uint32_t OFF_BARS = OFF_WPART + 4 * 2 * CO * 4;
  static constexpr uint32_t SMEM = OFF_BARS + 64 + 1024;
};

// layers 2-4.  in: RAW accumulators of the previous layer [B][Ci][Li] with in_stats [B][Ci] = (mean, rstd);
// out: RAW accumulators [B][CO][Lo] and part[b][co][tile] (FINAL: out = accumulators + bias, no partials).
__device__ __forceinline__ void builders_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 320 threads = 8 builder warps + a weight-producer warp + an MMA-issuer warp.
//   builders (thread = position tid & 127, half tid >> 7): stage the tile's input window (normalised), cut their
//     position's 16-tap im2col rows for two of the chunk's four channels, arrive on a_full[stage]; later the epilogue
//     (warp w drains TMEM lanes 32 (w & 3) and its half of the channels);
//   producer: requests a weight tile the moment the MMAs of its stage's previous user complete;
//   MMA issuer: waits for a_full + the weight tile, issues the 12 MMAs of a chunk, commits the stage free.
// Measured with clock64 stamps before this split (one 128-thread group doing everything, MMAs issued by warp 0 after a
// block barrier): a chunk took norm 700 + build 450-870 + barriers 300 + "issue 12 MMAs" 600-1500 cycles - tcgen05.mma
// issue blocks while the tensor pipe's queue is full, so the issuing warp sat there for the MMAs' own duration with the
// other warps parked at the next barrier: build and MMA were serialised, and no loader change could move the kernel.
template <int CO, bool FINAL>
__global__ void __launch_bounds__(320, 1) wav_conv_tc2_kernel(const float* __restrict__ in, const float2* __restrict__ in_stats,
                                                              const uint8_t* __restrict__ tape, const float* __restrict__ bias,
                                                              float* __restrict__ out, float2* __restrict__ part, int Ci, int Li,
                                                              int Lo, int n_tiles) {
  using L = Lay2<CO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* raw = reinterpret_cast<float*>(sm + L::OFF_RAW);
  float* wpart = reinterpret_cast<float*>(sm + L::OFF_WPART);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::OFF_BARS);       // w_full[2], empty[2], acc, a_full[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L::OFF_BARS + 56);
  enum { W_FULL0 = 0, EMPTY0 = 2, ACC = 4, A_FULL0 = 5 };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = tid & 127, half = (tid >> 7) & 1;
  const int b = blockIdx.y, tile = blockIdx.x, lo0 = tile * 128, lo = lo0 + pos;
  const bool valid = lo < Lo;
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bars[A_FULL0], 8);                     // one arrival per builder warp
    mbar_init(&bars[A_FULL0 + 1], 8);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<CO>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sm_s = smem_u32(sm), bars_s = smem_u32(bars);
  const int n_chunks = Ci / CI_PER_CHUNK;

  if (warp == 8) {
    // ================= weight producer ==============================================================
    if (lane == 0) {
      for (int c = 0; c < n_chunks; ++c) {
        const int s = c & 1;
        if (c >= 2) mbar_wait(&bars[EMPTY0 + s], ((c >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&bars[W_FULL0 + s], 2 * L::B_IMG);
        bulk_g2s(sm + s * L::STAGE + 2 * A_IMG, tape + (size_t)c * 2 * L::B_IMG, 2 * L::B_IMG, &bars[W_FULL0 + s]);
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer: whole warp in uniform control flow, one elected lane issues ========
    constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
    constexpr uint32_t idesc = idesc_bf16(128, CO, 0, 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c & 1;
      mbar_wait_s(bars_s + 8 * (A_FULL0 + s), (c >> 1) & 1);
      mbar_wait_s(bars_s + 8 * (W_FULL0 + s), (c >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t a_hi = desc_lo32(sm_s + s * L::STAGE, 16), a_lo = desc_lo32(sm_s + s * L::STAGE + A_IMG, 16);
      const uint32_t b_hi = desc_lo32(sm_s + s * L::STAGE + 2 * A_IMG, 16),
                     b_lo = desc_lo32(sm_s + s * L::STAGE + 2 * A_IMG + L::B_IMG, 16);
#pragma unroll
      for (uint32_t ks = 0; ks < 4; ++ks) {
        umma_bf16_split_elect(tm, a_hi + 2 * ks, DH, b_hi + 2 * ks, DH, idesc, (c > 0 || ks > 0) ? 1u : 0u);
        umma_bf16_split_elect(tm, a_lo + 2 * ks, DH, b_hi + 2 * ks, DH, idesc, 1u);
        umma_bf16_split_elect(tm, a_hi + 2 * ks, DH, b_lo + 2 * ks, DH, idesc, 1u);
      }
      umma_commit_s_elect(bars_s + 8 * (EMPTY0 + s));
    }
    umma_commit_s_elect(bars_s + 8 * ACC);
  } else {
    // ================= builders, then the epilogue ====================================================
    // Warp w builds tile positions 16 w .. 16 w + 15 from ITS OWN window of the input (samples 96 w .. 96 w + 105 of
    // each channel, staged in a private 112-float row): no block barrier in the loop, only __syncwarp, so the eight
    // warps drift apart and hide each other's latencies (two block barriers per chunk cost 300 of 1500 cycles and kept
    // the warps in lock step).  Adjacent warps re-read 10 of 106 samples.
    constexpr int WWIN = 15 * STRIDE + 16;            // 106
    constexpr int WROW = 108;                         // floats per staged channel row of a warp (keeps 2 CTAs per SM at CO = 64)
    float* wraw = raw + warp * (CI_PER_CHUNK * WROW);
    const int w0 = (lo0 + 16 * warp) * STRIDE, n_ok = min(WWIN, Li - w0);      // may be <= 0 past the end of the row
    const float* inb = in + (size_t)b * Ci * Li + w0;
    const float2* stb = in_stats + (size_t)b * Ci;
    float nx[CI_PER_CHUNK][4];
    auto load_raw = [&](int c) {                      // coalesced: consecutive lanes, consecutive samples
#pragma unroll
      for (int cil = 0; cil < CI_PER_CHUNK; ++cil) {
        const float* src = inb + (size_t)(c * CI_PER_CHUNK + cil) * Li;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = lane + 32 * i;
          nx[cil][i] = idx < n_ok ? __ldg(src + idx) : 0.f;
        }
      }
    };
    load_raw(0);
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c & 1;
      uint8_t* stage = sm + s * L::STAGE;
      // normalise (InstanceNorm of the previous layer) + LeakyReLU once per element, into the warp's staging rows
#pragma unroll
      for (int cil = 0; cil < CI_PER_CHUNK; ++cil) {
        const float2 st = __ldg(stb + c * CI_PER_CHUNK + cil);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = lane + 32 * i;
          if (idx < WROW) {
            // Samples past the end of the row (loaded as 0) only meet the zero weight of the 16th tap or belong to
            // positions >= Lo, whose accumulator rows are never read: any finite value will do, no select needed.
            const float v = (nx[cil][i] - st.x) * st.y;     // the reference's order: subtract the mean, then scale
            wraw[cil * WROW + idx] = fmaxf(v, 0.3f * v);
          }
        }
      }
      if (c + 1 < n_chunks) load_raw(c + 1);          // in flight during the im2col build
      if (c >= 2) {                                   // the MMAs that read this stage (chunk c-2) are done
        mbar_wait(&bars[EMPTY0 + s], ((c >> 1) - 1) & 1);
      }
      __syncwarp();                                   // the warp's staging rows are complete
#pragma unroll
      for (int i = 0; i < 4; ++i) {                   // 16 positions x 4 channels x 2 halves = 128 units of 8 samples
        const int u = lane + 32 * i, pl = u & 15, cil = (u >> 4) & 3, hh = u >> 6;
        const int row = 16 * warp + pl;
        const float2* wr = reinterpret_cast<const float2*>(wraw + cil * WROW + pl * STRIDE + 8 * hh);
        float v[8];                                   // rows of positions >= Lo hold finite leftovers: their outputs are masked
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 t = wr[q];
          v[2 * q] = t.x;
          v[2 * q + 1] = t.y;
        }
        uint4 hi, lw;
        hi.x = pack_hi_lo(v[0], v[1], &lw.x);
        hi.y = pack_hi_lo(v[2], v[3], &lw.y);
        hi.z = pack_hi_lo(v[4], v[5], &lw.z);
        hi.w = pack_hi_lo(v[6], v[7], &lw.w);
        const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + ((uint32_t)((cil * 2 + hh) ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(stage + off) = hi;
        *reinterpret_cast<uint4*>(stage + A_IMG + off) = lw;
      }
      fence_proxy_async_smem();
      __syncwarp();                                   // also: every lane is done with the staging rows
      if (lane == 0) mbar_arrive(&bars[A_FULL0 + s]);
    }
    mbar_wait(&bars[ACC], 0);
    __syncwarp();
    tc_fence_after_sync();
    // epilogue: lane = output position, column = output channel; the two halves split the channels
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* dst = out + (size_t)b * CO * Lo + lo;
#pragma unroll 1
    for (int c0 = half * (CO / 2); c0 < (half + 1) * (CO / 2); c0 += 16) {
      float v[32];
      tmem_ld16(taddr + c0, v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[(size_t)(c0 + j) * Lo] = FINAL ? v[j] + __ldg(bias + c0 + j) : v[j];
      }
      if (!FINAL) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = valid ? v[j] : 0.f;
          v[16 + j] = v[j] * v[j];
        }
        butterfly32(v, lane);                         // lane = (square? << 4) | channel of this group
        wpart[(warp & 3) * 2 * CO + (lane >> 4) * CO + c0 + (lane & 15)] = v[0];
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (!FINAL) {
    for (int i = tid; i < 2 * CO; i += 320) {
      const float t = (wpart[i] + wpart[2 * CO + i]) + (wpart[4 * CO + i] + wpart[6 * CO + i]);
      float* d = reinterpret_cast<float*>(part + ((size_t)b * CO + (i % CO)) * n_tiles + tile);
      d[i / CO] = t;
    }
  }
  if (warp == 0) tmem_dealloc<CO>(tmem);
}

// The LAST layer (Conv1d(128, 256, 15, 6), 217 -> 34 positions, bias, no norm after it) with the rows of a tile taken from
// SEVERAL clips: one clip has only 34 output positions, so a tile per clip left 73 % of every MMA's rows empty.  Tile row
// r = global position g0 + r, clip g / 34, position g % 34; a tile spans at most 5 clips, whose whole input rows (217
// samples per channel, padded to 218 floats so that every window start stays 8-byte aligned) are staged per chunk.
constexpr int P_LO = 34, P_LI = 217, P_ROW = 218, P_CLIPS = 5, P_CO = 256;
struct LayP {
  static constexpr uint32_t B_IMG = P_CO * 128;
  static constexpr uint32_t STAGE = 2 * A_IMG + 2 * B_IMG;
  static constexpr uint32_t OFF_RAW = 2 * STAGE;
  static constexpr uint32_t RAW_FLOATS = P_CLIPS * CI_PER_CHUNK * P_ROW;        // 4360
  static constexpr uint32_t OFF_ST = OFF_RAW + RAW_FLOATS * 4;                    // [5 clips][4 channels] (mean, rstd)
  static constexpr uint32_t OFF_BARS = OFF_ST + P_CLIPS * CI_PER_CHUNK * 8;
  static constexpr uint32_t SMEM = OFF_BARS + 64 + 1024;
};
static_assert(LayP::SMEM <= 232448, "shared memory budget of the packed last layer");

__global__ void __launch_bounds__(320, 1) wav_conv4_packed_kernel(const float* __restrict__ in, const float2* __restrict__ in_stats,
                                                                  const uint8_t* __restrict__ tape, const float* __restrict__ bias,
                                                                  float* __restrict__ out, int Ci, int n_rows) {
  using L = LayP;
  constexpr int CO = P_CO;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* raw = reinterpret_cast<float*>(sm + L::OFF_RAW);
  float2* st_s = reinterpret_cast<float2*>(sm + L::OFF_ST);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::OFF_BARS);       // w_full[2], empty[2], acc, a_full[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L::OFF_BARS + 56);
  enum { W_FULL0 = 0, EMPTY0 = 2, ACC = 4, A_FULL0 = 5 };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = tid & 127, half = (tid >> 7) & 1;
  const int g0 = blockIdx.x * 128, g = g0 + pos;
  const bool valid = g < n_rows;
  const int clip0 = g0 / P_LO;
  const int n_clips_total = n_rows / P_LO;
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bars[A_FULL0], 8);
    mbar_init(&bars[A_FULL0 + 1], 8);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<CO>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sm_s = smem_u32(sm), bars_s = smem_u32(bars);
  const int n_chunks = Ci / CI_PER_CHUNK;

  if (warp == 8) {
    if (lane == 0) {
      for (int c = 0; c < n_chunks; ++c) {
        const int s = c & 1;
        if (c >= 2) mbar_wait(&bars[EMPTY0 + s], ((c >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&bars[W_FULL0 + s], 2 * L::B_IMG);
        bulk_g2s(sm + s * L::STAGE + 2 * A_IMG, tape + (size_t)c * 2 * L::B_IMG, 2 * L::B_IMG, &bars[W_FULL0 + s]);
      }
    }
  } else if (warp == 9) {
    constexpr uint32_t DH = desc_hi32(1024, (uint32_t)SWZ_128B);
    constexpr uint32_t idesc = idesc_bf16(128, CO, 0, 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c & 1;
      mbar_wait_s(bars_s + 8 * (A_FULL0 + s), (c >> 1) & 1);
      mbar_wait_s(bars_s + 8 * (W_FULL0 + s), (c >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t a_hi = desc_lo32(sm_s + s * L::STAGE, 16), a_lo = desc_lo32(sm_s + s * L::STAGE + A_IMG, 16);
      const uint32_t b_hi = desc_lo32(sm_s + s * L::STAGE + 2 * A_IMG, 16),
                     b_lo = desc_lo32(sm_s + s * L::STAGE + 2 * A_IMG + L::B_IMG, 16);
#pragma unroll
      for (uint32_t ks = 0; ks < 4; ++ks) {
        umma_bf16_split_elect(tm, a_hi + 2 * ks, DH, b_hi + 2 * ks, DH, idesc, (c > 0 || ks > 0) ? 1u : 0u);
        umma_bf16_split_elect(tm, a_lo + 2 * ks, DH, b_hi + 2 * ks, DH, idesc, 1u);
        umma_bf16_split_elect(tm, a_hi + 2 * ks, DH, b_lo + 2 * ks, DH, idesc, 1u);
      }
      umma_commit_s_elect(bars_s + 8 * (EMPTY0 + s));
    }
    umma_commit_s_elect(bars_s + 8 * ACC);
  } else {
    const uint32_t row_off = (uint32_t)(pos >> 3) * 1024u + (uint32_t)(pos & 7) * 128u;
    // staged elements of this thread: e = tid + 256 i over [clip j][channel cil][sample k < 217]; chunk-invariant parts
    constexpr int N_E = P_CLIPS * CI_PER_CHUNK * P_LI;          // 4340
    constexpr int PER_T = (N_E + 255) / 256;                    // 17
    int goff[PER_T];                 // global offset without the chunk's channel base, or -1
    short soff[PER_T], sidx[PER_T];  // staging offset; (j * 4 + cil) for the statistics
#pragma unroll
    for (int i = 0; i < PER_T; ++i) {
      const int e = tid + 256 * i;
      const int jc = e / P_LI, k = e - jc * P_LI, j = jc >> 2, cil = jc & 3;
      const bool ok = e < N_E && clip0 + j < n_clips_total;
      goff[i] = ok ? ((clip0 + j) * Ci + cil) * P_LI + k : -1;
      soff[i] = (short)(e < N_E ? jc * P_ROW + k : -1);
      sidx[i] = (short)jc;
    }
    float nx[PER_T];
    auto load_raw = [&](int c) {
#pragma unroll
      for (int i = 0; i < PER_T; ++i) nx[i] = goff[i] >= 0 ? __ldg(in + goff[i] + c * CI_PER_CHUNK * P_LI) : 0.f;
    };
    auto load_stats = [&](int c) {   // (mean, rstd) of the chunk's 4 channels for the tile's clips
      if (tid < P_CLIPS * CI_PER_CHUNK) {
        const int j = tid >> 2, cil = tid & 3;
        st_s[tid] = clip0 + j < n_clips_total ? __ldg(in_stats + (size_t)(clip0 + j) * Ci + c * CI_PER_CHUNK + cil)
                                              : make_float2(0.f, 1.f);
      }
    };
    const int my_j = valid ? g / P_LO - clip0 : 0, my_lo = valid ? g % P_LO : 0;
    load_raw(0);
    load_stats(0);
    builders_bar();
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c & 1;
      uint8_t* stage = sm + s * L::STAGE;
#pragma unroll
      for (int i = 0; i < PER_T; ++i) {
        if (soff[i] >= 0) {
          const float2 st = st_s[sidx[i]];
          float v = (nx[i] - st.x) * st.y;
          v = v > 0.f ? v : 0.3f * v;
          raw[soff[i]] = goff[i] >= 0 ? v : 0.f;
        }
      }
      if (tid < P_CLIPS * CI_PER_CHUNK) raw[tid * P_ROW + P_LI] = 0.f;       // the pad sample (zero-weight 16th tap of the last window)
      if (c + 1 < n_chunks) load_raw(c + 1);
      if (c >= 2) mbar_wait(&bars[EMPTY0 + s], ((c >> 1) - 1) & 1);
      builders_bar();                                 // staging rows complete; everybody has read this chunk's statistics
      if (c + 1 < n_chunks) load_stats(c + 1);
#pragma unroll
      for (int cc = 0; cc < CI_PER_CHUNK / 2; ++cc) {
        const int cil = 2 * half + cc;
        const float2* wr = reinterpret_cast<const float2*>(raw + (my_j * CI_PER_CHUNK + cil) * P_ROW + my_lo * STRIDE);
        float v[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 t = valid ? wr[i] : make_float2(0.f, 0.f);
          v[2 * i] = t.x;
          v[2 * i + 1] = t.y;
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint4 hi, lw;
          hi.x = pack_hi_lo(v[8 * hh + 0], v[8 * hh + 1], &lw.x);
          hi.y = pack_hi_lo(v[8 * hh + 2], v[8 * hh + 3], &lw.y);
          hi.z = pack_hi_lo(v[8 * hh + 4], v[8 * hh + 5], &lw.z);
          hi.w = pack_hi_lo(v[8 * hh + 6], v[8 * hh + 7], &lw.w);
          const uint32_t off = row_off + ((uint32_t)((cil * 2 + hh) ^ (pos & 7)) << 4);
          *reinterpret_cast<uint4*>(stage + off) = hi;
          *reinterpret_cast<uint4*>(stage + A_IMG + off) = lw;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[A_FULL0 + s]);
      builders_bar();                                 // every builder is done with the staging rows (and st_s is rewritten)
    }
    mbar_wait(&bars[ACC], 0);
    __syncwarp();
    tc_fence_after_sync();
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* dst = out + ((size_t)(clip0 + my_j) * CO) * P_LO + my_lo;
#pragma unroll 1
    for (int c0 = half * (CO / 2); c0 < (half + 1) * (CO / 2); c0 += 16) {
      float v[16];
      tmem_ld16(taddr + c0, v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[(size_t)(c0 + j) * P_LO] = v[j] + __ldg(bias + c0 + j);
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<CO>(tmem);
}

template <int CO, bool FINAL>
int launch_conv2(ls_handle* h, const float* in, const float2* st, const uint8_t* tape, const float* bias, float* out,
                 float2* part, int nb, int Ci, int Li, int Lo, cudaStream_t s) {
  const int n_tiles = (Lo + 127) / 128;
  wav_conv_tc2_kernel<CO, FINAL><<<dim3(n_tiles, nb), 320, Lay2<CO>::SMEM, s>>>(in, st, tape, bias, out, part, Ci, Li, Lo, n_tiles);
  LS_LAUNCH_CHECK(h);
  return LS_OK;
}

}  // namespace

// The whole encoder with the InstanceNorm passes fused away (tensor-core path).  audio [nb][L0] -> out_cm [nb][256][34];
// wav_a / wav_b: the handle's ping-pong buffers; w0: layer-1 weights [32][1][15]; b3: the last layer's bias.
int lsw_encoder_fused(ls_handle* h, const float* audio, const float* /*w0: uploaded by lsw_init*/, const float* b3, float* out_cm,
                      int nb, int L0, int L1, int L2, int L3, int L4, cudaStream_t s) {
  WavTc* wt = static_cast<WavTc*>(h->wavtc);
  if (!wt) return ls_fail(h, LS_EUNSUPPORTED, "tensor-core WavEncoder not initialised");
  const int t1 = (L1 + C1_SPAN - 1) / C1_SPAN, t2 = (L2 + 127) / 128, t3 = (L3 + 127) / 128;
  const size_t need_part = (size_t)nb * std::max(std::max(32 * t1, 64 * t2), 128 * t3), need_stats = (size_t)nb * 128;
  if (wt->v2_part_elems < need_part || wt->v2_stats_elems < need_stats) {
    if (wt->v2_part) cudaFree(wt->v2_part);
    if (wt->v2_stats) cudaFree(wt->v2_stats);
    wt->v2_part = nullptr; wt->v2_stats = nullptr; wt->v2_part_elems = wt->v2_stats_elems = 0;
    if (cudaMalloc(&wt->v2_part, need_part * sizeof(float2)) != cudaSuccess ||
        cudaMalloc(&wt->v2_stats, need_stats * sizeof(float2)) != cudaSuccess)
      return ls_fail(h, LS_ENOMEM, "WavEncoder statistics buffers");
    wt->v2_part_elems = need_part;
    wt->v2_stats_elems = need_stats;
  }
  if (!wt->v2_attr_done) {
    LS_CUDA(h, cudaFuncSetAttribute(wav_conv_tc2_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay2<64>::SMEM));
    LS_CUDA(h, cudaFuncSetAttribute(wav_conv_tc2_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay2<128>::SMEM));
    LS_CUDA(h, cudaFuncSetAttribute(wav_conv_tc2_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay2<256>::SMEM));
    LS_CUDA(h, cudaFuncSetAttribute(wav_conv4_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LayP::SMEM));
    wt->v2_attr_done = true;
  }
  float2* part = static_cast<float2*>(wt->v2_part);
  float2* stats = static_cast<float2*>(wt->v2_stats);
  int rc;
  Conv1W cw;
  memcpy(cw.w, wt->w1_host, sizeof(cw.w));
  wav_conv1_kernel<<<dim3(t1, nb), C1_TILE, 0, s>>>(audio, cw, h->wav_a, part, L0, L1, t1);
  LS_LAUNCH_CHECK(h);
  in_finalize_kernel<<<(nb * 32 + 255) / 256, 256, 0, s>>>(part, nb * 32, t1, L1, stats);
  LS_LAUNCH_CHECK(h);
  if ((rc = launch_conv2<64, false>(h, h->wav_a, stats, wt->tape[0], nullptr, h->wav_b, part, nb, 32, L1, L2, s))) return rc;
  in_finalize_kernel<<<(nb * 64 + 255) / 256, 256, 0, s>>>(part, nb * 64, t2, L2, stats);
  LS_LAUNCH_CHECK(h);
  if ((rc = launch_conv2<128, false>(h, h->wav_b, stats, wt->tape[1], nullptr, h->wav_a, part, nb, 64, L2, L3, s))) return rc;
  in_finalize_kernel<<<(nb * 128 + 255) / 256, 256, 0, s>>>(part, nb * 128, t3, L3, stats);
  LS_LAUNCH_CHECK(h);
  if (L3 == P_LI && L4 == P_LO) {                     // the shipped geometry: tiles packed across clips
    wav_conv4_packed_kernel<<<(nb * P_LO + 127) / 128, 320, LayP::SMEM, s>>>(h->wav_a, stats, wt->tape[2], b3, out_cm, 128, nb * P_LO);
    LS_LAUNCH_CHECK(h);
    return LS_OK;
  }
  return launch_conv2<256, true>(h, h->wav_a, stats, wt->tape[2], b3, out_cm, nullptr, nb, 128, L3, L4, s);
}
