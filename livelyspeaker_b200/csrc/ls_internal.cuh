// Internal definitions shared by the translation units of libls_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/livelyspeaker_b200.h"

constexpr int LS_D = 512;          // latent_dim (scripts/mdm_utils/parser_util.py: --latent_dim 512)
constexpr int LS_F = 34;           // frames per clip
constexpr int LS_AF = 256;         // WavEncoder output channels (audio_enc.py:18)
constexpr int LS_SPK = 256;        // speaker embedding width (RAG.py:66)
constexpr int LS_NPRE = 4;         // n_pre_seq (RAG.py:70)
constexpr int LS_MAX_LAYERS = 16;
constexpr int LS_MAX_S = 36;       // tokens per pass: 34 frames + style (+ emotion)

// Device-side weight views handed to kernels by value.
struct LsLayerW {
  const float* ln1_a; const float* ln1_b;   // block1.0.alpha/beta [512]
  const float* w_tok; const float* b_tok;   // block1.1.weight [S,S], bias [S]
  const float* ln2_a; const float* ln2_b;   // block2.0.alpha/beta [512]
  const float* w_ch_t;                      // block2.1.weight transposed: [k][c]
  const float* b_ch;                        // [512]
};

struct LsWeights {
  LsLayerW layer[LS_MAX_LAYERS];
  const float* w_x_t;     // input_mapping.weight[:, 0:JD]^T          [JD][512]
  const float* w_o_t;     // input_mapping.weight[:, JD:2JD]^T        [JD][512]
  const float* w_bit;     // input_mapping.weight[:, 2JD]             [512]
  const float* w_a_t;     // input_mapping.weight[:, 2JD+1:]^T        [256][512]
  const float* b_in;      // input_mapping.bias                       [512]
  const float* w_out;     // output_process.poseFinal.weight          [JD][512]
  const float* b_out;     //                                          [JD]
  const float* spk_emb;   // speaker_embedding.weight                 [n_spk][256]
  const float* w_mu_t; const float* b_mu;   // speaker_mu   ^T [256][512], [512]
  const float* w_lv_t; const float* b_lv;   // speaker_logvar
  const float* emo_emb;   // emotion_embedding.weight [n_emo][512] or nullptr
  const float* emb_table; // time_embed(pe[t])  [max_timestep][512]
};

struct RawTensor {
  std::string key;
  std::vector<int64_t> shape;
  float* dev = nullptr;
  int64_t numel = 0;
  bool required = true;
  bool loaded = false;
};

struct ls_handle {
  ls_config cfg{};
  int JD = 0, S = 0;
  std::string err;
  std::vector<RawTensor> raw;       // reference-named fp32 tensors
  std::vector<void*> allocs;        // everything cudaMalloc'ed by this handle
  LsWeights w{};
  bool finalized = false;
  int cond_batch = 0;               // batch of the last ls_precompute_cond (0 = none)
  int impl = LS_IMPL_AUTO;
  int64_t launches = 0;
  // step-invariant conditioning (sized for max_batch)
  float* A = nullptr;       // [B,34,512] audio half of input_mapping
  float* P = nullptr;       // [B,34,512] prefix half + bit + bias
  float* z_mu = nullptr;    // [B,512]
  float* z_lv = nullptr;    // [B,512]
  float* emo_tok = nullptr; // [B,512] (BEAT)
  // scratch
  float* out_c = nullptr;   // [B,JD,34] cond denoiser output
  float* out_u = nullptr;   // [B,JD,34] uncond
  float* wav_a = nullptr;   // WavEncoder ping
  float* wav_b = nullptr;   // WavEncoder pong
  float* af = nullptr;      // [chunk,256,34] encoder output (channel-major)
  int64_t* t_tmp = nullptr; // [B] timesteps for batch-uniform calls
  float* w1_t = nullptr;    // time_embed.0.weight^T
  float* w2_t = nullptr;    // time_embed.2.weight^T
  int wav_chunk = 0;
  float* dbg_h = nullptr;   // ls_debug_hidden target (nullptr = off)
  int dbg_layer = -2;
  // ls_cfg_forward_grad / ls_cfg_backward (allocated on first use, grown with the batch)
  float* grad_ckpt = nullptr;   // [B][2][n_layers][S][512]
  float* grad_gx = nullptr;     // [2][B][JD][34]
  int64_t* grad_t = nullptr;    // [B] timesteps of the saved forward
  int grad_cap = 0, grad_batch = 0;
  void* fused = nullptr;    // state of the tcgen05 path (ls_fused.cu)
  void* wavtc = nullptr;    // state of the tcgen05 WavEncoder convolutions (ls_wavenc_tc.cu)
};

int ls_fail(ls_handle* h, int code, const char* fmt, ...);

#define LS_CUDA(h, expr)                                                              \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return ls_fail((h), LS_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,        \
                     cudaGetErrorString(e__));                                        \
  } while (0)

#define LS_LAUNCH_CHECK(h)                                                            \
  do {                                                                                \
    (h)->launches++;                                                                  \
    LS_CUDA((h), cudaGetLastError());                                                 \
  } while (0)

// ---- kernels implemented across the .cu files -------------------------------------
// ls_precompute.cu
int lsk_wav_encoder(ls_handle* h, int B, const float* audio, float* out_cm, cudaStream_t s);  // out [B,256,34]
int lsk_cond_proj(ls_handle* h, int B, int b0, const float* af_cm, float* origin_x, const int64_t* vid,
                  const int64_t* emo, int64_t emo_stride, int mutate_origin, cudaStream_t s);
int lsk_time_embed_table(ls_handle* h, const float* pe, const float* w1, const float* b1, const float* w2,
                         const float* b2, float* table, int n_t, cudaStream_t s);
int lsk_transpose(ls_handle* h, const float* in, float* out, int rows, int cols, int ld_in, cudaStream_t s);
int lsk_cm_to_fm(ls_handle* h, const float* in_cm, float* out_fm, int B, cudaStream_t s);
// ls_denoise_simt.cu
// ckpt (or nullptr): [B][2 passes][n_layers][S][512] inputs of every MLPblock, for lsk_denoise_simt_bwd
int lsk_denoise_simt(ls_handle* h, int B, const float* x, const int64_t* t, int t_uniform, int pass_mask,
                     const float* eps_c, const float* eps_u, float* out_c, float* out_u, cudaStream_t s,
                     float* ckpt = nullptr, const uint8_t* cond_drop = nullptr);
int lsk_denoise_simt_bwd(ls_handle* h, int B, const int64_t* t, const float* ckpt, const float* grad_out,
                         const float* scale, float* gx, cudaStream_t s);
// ls_update.cu
int lsk_cfg_combine(ls_handle* h, int B, const float* out_c, const float* out_u, const float* scale,
                    float* out, cudaStream_t s);
int lsk_cfg_update(ls_handle* h, int B, const ls_step_params* p, const float* out_c, const float* out_u,
                   const float* scale, const float* x_t, const float* noise, int64_t sb, int64_t sj, int64_t sf,
                   float* x_prev, float* pred_x0, cudaStream_t s);
int lsk_axpby(ls_handle* h, int64_t n, const float* a, const float* b, float ca, float cb, float* out,
              cudaStream_t s);
int lsk_pass_scale(ls_handle* h, int B, int cond, const uint8_t* drop, float* scale, cudaStream_t s);
// ls_wavenc_tc.cu
int lsw_init(ls_handle* h, const float* const w[3], cudaStream_t s, const float* w0 = nullptr);   // w0: layer-1 weights
void lsw_destroy(ls_handle* h);
int lsw_available(const ls_handle* h);
int lsw_conv(ls_handle* h, int layer, const float* in, const float* bias, float* out, int nb, int Li, int Lo,
             cudaStream_t s);
int lsw_audio_proj(ls_handle* h, const float* af_cm, int nb, float* A, cudaStream_t s);
int lsw_encoder_fused(ls_handle* h, const float* audio, const float* w0, const float* b3, float* out_cm, int nb, int L0,
                      int L1, int L2, int L3, int L4, cudaStream_t s);   // whole WavEncoder, InstanceNorm fused into the convs   // A [nb*34][512] from af [nb][256][34]
// ls_fused.cu
int lsf_init(ls_handle* h, cudaStream_t s);            // build bf16 weight tapes; 0 if available
void lsf_destroy(ls_handle* h);
int lsf_available(const ls_handle* h);
// t_dev: per-clip ORIGINAL timesteps on the device (then p->t_model is ignored; mode 2 callers) or nullptr
int lsf_steps(ls_handle* h, int B, int n_steps, const ls_step_params* p, const ls_step_io* io, int precise,
              const float* x_in, const float* scale, cudaStream_t s, const int64_t* t_dev = nullptr);
