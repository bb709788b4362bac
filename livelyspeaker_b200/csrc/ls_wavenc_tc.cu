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
    delete w;
    h->wavtc = nullptr;
  }
}

// w[i]: conv weights of layers 2..4 ([Co, Ci, 15]); builds the three tapes
int lsw_init(ls_handle* h, const float* const w[3], cudaStream_t s) {
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
