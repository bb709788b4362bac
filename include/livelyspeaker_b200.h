/*
 * livelyspeaker_b200.h - C ABI of the B200-native RAG sampling path.
 *
 * The reference (zyhbili/LivelySpeaker @ 7f6ccd1) is pure Python/PyTorch and has
 * no FFI of its own; this header is therefore the boundary a
 * maintainer would bind with ctypes from the reference's Python call sites.
 * Each entry point names the reference code it replaces.  INTEGRATION.md shows
 * the ctypes stubs.
 *
 * Conventions
 *   - every function returns 0 on success, a negative LS_E* code on failure;
 *     ls_last_error() gives the message.  No exceptions cross the ABI.
 *   - all tensor pointers are DEVICE pointers on the handle's device unless the
 *     parameter name ends in _host; fp32 unless stated; layouts are the
 *     reference's logical layouts, dense, row-major unless strides are passed.
 *   - calls are asynchronous on the given stream (a cudaStream_t passed as
 *     void*; NULL = legacy default stream).  No allocation happens after
 *     ls_create(): workspaces sized for cfg.max_batch live in the handle.
 *   - one handle per GPU per model; a handle is not thread-safe.
 */
#ifndef LIVELYSPEAKER_B200_H
#define LIVELYSPEAKER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LS_ABI_VERSION 1

enum {
  LS_OK = 0,
  LS_EINVAL = -1,   /* bad argument / shape mismatch            */
  LS_ESTATE = -2,   /* call order violated (weights, cond, ...) */
  LS_ECUDA = -3,    /* a CUDA runtime call failed               */
  LS_ENOMEM = -4,
  LS_EUNSUPPORTED = -5
};

typedef struct ls_handle ls_handle;

/* Model geometry.  TED: njoints 9, nfeats 3, n_pre_emb 1, audio_len 36267.
 * BEAT: 47, 6, 2 (style + emotion token), 36266.
 * scripts/mdm_utils/model_util.py:20-37, scripts/model/RAG.py:17-74,
 * scripts_beat/model/RAG.py:56,72-74. */
typedef struct ls_config {
  int32_t njoints;
  int32_t nfeats;
  int32_t n_frames;      /* must be 34: fixed by Conv1d(S,S,1) + WavEncoder strides */
  int32_t n_pre_emb;     /* 1 (TED) or 2 (BEAT)                                     */
  int32_t latent_dim;    /* must be 512 in this build                               */
  int32_t n_layers;      /* MLPblock count, <= 16                                   */
  int32_t audio_len;     /* samples per clip; WavEncoder must map it to n_frames    */
  int32_t n_speakers;    /* rows of speaker_embedding (1400)                        */
  int32_t n_emotions;    /* rows of emotion_embedding (8), 0 for TED                */
  int32_t max_batch;     /* workspaces are sized for this many clips                */
  int32_t max_timestep;  /* time-embedding table covers original t in [0, this)     */
  int32_t device;        /* CUDA device ordinal                                     */
} ls_config;

/* Sampler update applied after the guided x0 prediction.
 *   mode 0, ancestral  (gaussian_diffusion.py:260-282, 507-558):
 *       x_prev = c[0]*x0 + c[1]*x_t + add_noise * exp(0.5*c[2]) * noise
 *       c = { posterior_mean_coef1[i], posterior_mean_coef2[i],
 *             posterior_log_variance_clipped[i] }               (fp64 -> fp32 casts)
 *   mode 1, DDIM       (gaussian_diffusion.py:418-422, 745-798):
 *       eps    = (c[0]*x_t - x0) / c[1]
 *       x_prev = x0*c[2] + c[3]*eps + add_noise * c[4] * noise
 *       c = { sqrt_recip_alphas_cumprod[i], sqrt_recipm1_alphas_cumprod[i],
 *             sqrt(abar_prev), sqrt(1-abar_prev-sigma^2), sigma }  (fp32)
 *   mode 2: no update, only pred_x0 is written (p_mean_variance callers). */
typedef struct ls_step_params {
  int32_t mode;
  int32_t t_model;        /* ORIGINAL timestep fed to the denoiser = timestep_map[i]
                             (respace.py:118-130)                                     */
  int32_t clip_denoised;  /* clamp x0 to [-1,1] (gaussian_diffusion.py:365-371)       */
  int32_t add_noise;      /* (i != 0)                                                 */
  float c[8];
} ls_step_params;

/* Per-step buffers of ls_step_multi: the draws and outputs of ONE step, same meaning as
 * the ls_step arguments of the same names.                                              */
typedef struct ls_step_io {
  const float* eps_cond;     /* [B,512] randn of the cond pass' reparameterize           */
  const float* eps_uncond;   /* [B,512] ... of the uncond pass                            */
  const float* noise;        /* step noise, strided (may be NULL when add_noise == 0)     */
  int64_t noise_sb, noise_sj, noise_sf;
  float* x_prev;             /* [B,J*D,F] sample after this step                          */
  float* pred_x0;            /* [B,J*D,F] or NULL                                         */
} ls_step_io;

#define LS_MAX_FUSED_STEPS 16

/* Which implementation ls_step / ls_cfg_forward use for the denoiser. */
enum {
  LS_IMPL_AUTO = 0,       /* fused tcgen05 kernel when built in, else simt            */
  LS_IMPL_SIMT = 1,       /* fp32 CUDA-core kernels (exact-order reference on device) */
  LS_IMPL_TC_BF16X3 = 2,  /* tcgen05, bf16 hi/lo split operands, fp32 accumulate      */
  LS_IMPL_TC_BF16 = 3     /* tcgen05, plain bf16 operands (fast mode, looser parity)  */
};

int ls_abi_version(void);
const char* ls_last_error(const ls_handle* h); /* h may be NULL: error of the last failed ls_create */

/* RAG.__init__ (scripts/model/RAG.py:17-74): allocates weights + workspaces. */
int ls_create(ls_handle** out, const ls_config* cfg);
void ls_destroy(ls_handle* h);

/* load_model_wo_clip / load_state_dict (scripts/mdm_utils/model_util.py:5-10):
 * one call per state_dict entry, reference key names and shapes, fp32, device
 * memory.  Unknown keys -> LS_EINVAL (the reference asserts no unexpected keys);
 * `clip_model.*` and the three `pe` buffers are accepted and ignored/used.       */
int ls_load_weight(ls_handle* h, const char* key, const float* dev_ptr,
                   const int64_t* shape, int32_t ndim, void* stream);
/* Checks that every required key arrived, derives the kernel-side layouts
 * (transposed / bf16 hi+lo / padded copies) and the time-embedding table
 * emb[t] = time_embed(pe[t]) (scripts/model/mlp_module.py:123-136).              */
int ls_finalize_weights(ls_handle* h, void* stream);

int ls_set_impl(ls_handle* h, int32_t impl);
int ls_get_impl(const ls_handle* h);   /* resolved implementation (never AUTO) */

/* Everything in RAG.forward that does not depend on x_t / t
 * (scripts/model/RAG.py:106-120): WavEncoder (audio_enc.py:6-25), the audio /
 * prefix halves of input_mapping, speaker mu / logvar, BEAT emotion token.
 * origin_x: [B, J*D, F]; frames >= 4 are ignored and, when mutate_origin != 0,
 * zeroed in place like RAG.py:110 does to the caller's tensor.
 * emo may be NULL (TED); emo_stride is the element stride between clips
 * (the reference reads y['emo'][:, 0]).                                          */
int ls_precompute_cond(ls_handle* h, int32_t B, const float* audio, float* origin_x,
                       const int64_t* vid_indices, const int64_t* emo, int64_t emo_stride,
                       int32_t mutate_origin, void* stream);

/* WavEncoder.forward alone (scripts/model/audio_enc.py:22-25): [B,L] -> [B,34,256]. */
int ls_wav_encoder(ls_handle* h, int32_t B, const float* audio, float* out, void* stream);

/* RAG.forward (scripts/model/RAG.py:98-133) on the cond set up by
 * ls_precompute_cond.  t: ORIGINAL timesteps, one per clip.  uncond != 0 zeroes the
 * audio embedding (mask_cond force_mask, RAG.py:80-83).  style_eps [B,512] is the
 * randn of reparameterize (RAG.py:10-13).  out [B,J*D,F]; z_mu / z_logvar [B,512]
 * may be NULL.  On the tensor-core implementations this is the fused kernel
 * (which always runs both guidance passes of a clip) with a per-clip combination
 * weight of 1 (cond) / 0 (uncond); LS_IMPL_SIMT runs the single fp32 pass.          */
int ls_model_forward(ls_handle* h, int32_t B, const float* x, const int64_t* t,
                     int32_t uncond, const float* style_eps, float* out,
                     float* z_mu, float* z_logvar, void* stream);

/* RAG.forward in TRAINING mode (scripts/model/RAG.py:80-96, 98-133): like ls_model_forward, with the per-clip condition
 * dropout of mask_cond - cond_drop [B] bytes on the device, 1 = this clip's audio embedding is replaced by zeros (the
 * caller draws the Bernoulli mask, keeping torch's generator order), NULL = no dropout.  Exact-order fp32 path.
 * Forward only: gradients with respect to the weights are not built.                           */
int ls_model_forward_train(ls_handle* h, int32_t B, const float* x, const int64_t* t,
                           const uint8_t* cond_drop, const float* style_eps, float* out,
                           float* z_mu, float* z_logvar, void* stream);

/* ClassifierFreeSampleModel.forward (scripts/model/cfg_sampler.py:24-31):
 * out = out_u + scale[b] * (out_c - out_u).                                       */
int ls_cfg_forward(ls_handle* h, int32_t B, const float* x, const int64_t* t,
                   const float* eps_cond, const float* eps_uncond, const float* scale,
                   float* out, void* stream);

/* The differentiable model call of the *_with_grad samplers (gaussian_diffusion.py:444-505 condition_*_with_grad,
 * 560-606 p_sample_with_grad, 800-855 ddim_sample_with_grad): there the reference runs p_mean_variance under
 * torch.enable_grad() with x.requires_grad_(), so that cond_fn(x, t, p_mean_var, **kwargs) can differentiate a function of
 * p_mean_var['pred_xstart'] with respect to x.
 *   ls_cfg_forward_grad: ls_cfg_forward on the exact-order fp32 path that also keeps the input of every MLPblock
 *     (B * 2 * n_layers * S * 512 floats; allocated on first use and grown with the batch - the one exception to
 *     "no allocation after ls_create").
 *   ls_cfg_backward: grad_x [B,J*D,F] = J^T grad_out for J = d out / d x of the LAST ls_cfg_forward_grad call of this
 *     handle (same B; scale = the guidance scales of that call).  The conditioning, the style draws and the weights are
 *     constants of this derivative; clamping / denoised_fn are the caller's (torch autograd's) business.               */
int ls_cfg_forward_grad(ls_handle* h, int32_t B, const float* x, const int64_t* t,
                        const float* eps_cond, const float* eps_uncond, const float* scale,
                        float* out, void* stream);
int ls_cfg_backward(ls_handle* h, int32_t B, const float* grad_out, const float* scale,
                    float* grad_x, void* stream);

/* One denoising step with a batch-uniform timestep: _WrappedModel +
 * ClassifierFreeSampleModel + RAG + p_mean_variance + p_sample / ddim_sample
 * (respace.py:118-130, cfg_sampler.py:24-31, RAG.py:98-133,
 * gaussian_diffusion.py:284-399, 507-558, 745-798).
 * noise is addressed as noise[b*noise_sb + jd*noise_sj + f*noise_sf] so the
 * reference's memory-order draw (DESIGN.md "RNG layout") can be passed as is;
 * noise_sb = 0 implements const_noise.  x_prev / pred_x0 are dense [B,J*D,F];
 * x_prev may alias x_t; pred_x0 may be NULL.                                     */
int ls_step(ls_handle* h, int32_t B, const ls_step_params* p, const float* x_t,
            const float* eps_cond, const float* eps_uncond, const float* noise,
            int64_t noise_sb, int64_t noise_sj, int64_t noise_sf, const float* scale,
            float* x_prev, float* pred_x0, void* stream);

/* n_steps (<= LS_MAX_FUSED_STEPS) consecutive iterations of the p_sample_loop /
 * ddim_sample_loop body (gaussian_diffusion.py:718-743, 986-1014) in one launch.
 * p and io are HOST arrays of n_steps entries in execution order; step k reads
 * x_t (k = 0) or io[k-1].x_prev, so the x_prev buffers must be distinct and must
 * not alias x_t when n_steps > 1.  Results equal n_steps calls of ls_step bit for
 * bit; the fused kernel schedules (clip, step) items across the SMs, which removes
 * the tail of a 512-clip batch on 148 SMs.  All mode fields must be 0 or 1.       */
int ls_step_multi(ls_handle* h, int32_t B, int32_t n_steps, const ls_step_params* p,
                  const ls_step_io* io, const float* x_t, const float* scale, void* stream);

/* q_sample (gaussian_diffusion.py:240-258): out = c_x0*x0 + c_noise*noise, n elements. */
int ls_q_sample(ls_handle* h, int64_t n, const float* x0, const float* noise,
                float c_x0, float c_noise, float* out, void* stream);

/* ---- SAG decoder ------------------------------------------------
 * Decoder_TRANSFORMER.forward (scripts/model/motionclip_module.py:137-183): the CLIP text
 * feature z -> coarse motion that LivelySpeaker sampling uses as init_image
 * (scripts/test_LivelySpeaker_ted.py:80-113).  Stateless: the caller passes device pointers
 * to the weights, the matrices TRANSPOSED ([in][out], contiguous) so that the kernel reads
 * them coalesced; biases / LayerNorm vectors as in the state_dict.                      */
#define LS_SAG_MAX_LAYERS 8
typedef struct ls_sag_layer {
  const float* sa_in_wt;   /* self_attn.in_proj_weight^T        [512][1536] */
  const float* sa_in_b;    /* self_attn.in_proj_bias            [1536]      */
  const float* sa_out_wt;  /* self_attn.out_proj.weight^T       [512][512]  */
  const float* sa_out_b;
  const float* ca_v_wt;    /* multihead_attn.in_proj_weight[1024:1536]^T [512][512] (the memory is one token) */
  const float* ca_v_b;     /* multihead_attn.in_proj_bias[1024:1536]               */
  const float* ca_out_wt;  /* multihead_attn.out_proj.weight^T  [512][512]  */
  const float* ca_out_b;
  const float* l1_wt;      /* linear1.weight^T                  [512][1024] */
  const float* l1_b;
  const float* l2_wt;      /* linear2.weight^T                  [1024][512] */
  const float* l2_b;
  const float *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b;   /* norm1..3 weight / bias [512] */
} ls_sag_layer;
typedef struct ls_sag_weights {
  int32_t n_layers, njoints, nfeats, n_frames, n_pre_poses, latent_dim, ff_size, n_heads;
  const float* map_wt;     /* mapping.weight^T                  [J*D+1][512] */
  const float* map_b;
  const float* pe;         /* sequence_pos_encoder.pe rows 0..n_frames-1, row stride pe_stride floats */
  int64_t pe_stride;
  const float* fin_wt;     /* finallayer.weight^T               [512][J*D]  */
  const float* fin_b;
  ls_sag_layer layer[LS_SAG_MAX_LAYERS];
} ls_sag_weights;
/* x [B,J*D,F] (frames >= n_pre_poses are ignored), z [B,512], mask [B,F] bytes (0 = padded,
 * may be NULL) -> out [B,J*D,F].  Errors: ls_last_error(NULL).                           */
int ls_sag_decode(const ls_sag_weights* w, int32_t B, const float* x, const float* z,
                  const uint8_t* mask, float* out, void* stream);
/* The same decoder on the tensor cores (tcgen05, bf16x3 split operands, fp32 accumulation; self-attention, softmax,
 * LayerNorm and GELU in fp32): ls_sag_create builds the weight tapes of the four projections per layer from *w and
 * folds each layer's single-token cross-attention into one matrix (those matrices are read during the call only);
 * biases / LayerNorm vectors / mapping / finallayer / pe are read by every decode and must stay valid for the handle's
 * life.  Workspaces for max_batch clips (557 KB per clip) live in the handle.  ls_sag_decode_tc has ls_sag_decode's
 * contract; ls_sag_decode stays the exact-order fp32 cross-check.  ls_sag_launch_count: kernels launched so far.     */
typedef struct ls_sag ls_sag;
int ls_sag_create(ls_sag** out, const ls_sag_weights* w, int32_t max_batch, int32_t device, void* stream);
int ls_sag_decode_tc(ls_sag* s, int32_t B, const float* x, const float* z, const uint8_t* mask,
                     float* out, void* stream);
int64_t ls_sag_launch_count(const ls_sag* s);
void ls_sag_destroy(ls_sag* s);

/* The n torch.randn / randn_like draws of a fused chunk in one launch: tensor i (dense, fp32,
 * numels[i] elements, written in memory order) receives exactly the values torch's CUDA
 * generator with `seed` would produce for it at Philox offset `philox_offset` + the increments
 * of tensors 0..i-1 (ATen/native/cuda/DistributionTemplates.h).  *total_increment is what the
 * caller must add to the generator's offset afterwards.  outs / numels are HOST arrays.  The
 * Python layer verifies the equality against torch once per process before using it.        */
#define LS_RANDN_MAX_TENSORS 48
int ls_randn_torch_compat(int32_t n, float* const* outs, const int64_t* numels, uint64_t seed,
                          uint64_t philox_offset, uint64_t* total_increment, int32_t device,
                          void* stream);

/* ---- post-sampling rhythm metric --------------------------------------
 * scripts/test_RAG_ted.py:84-111: the sampler output [B, njoints*3, F] (device) -> per-frame angle change of the listed
 * joint pairs (angle_diff [B,F], column 0 = 0) and the motion-beat mask (beat_mask [B,F] bytes: frame t in [2, F-1) is a
 * beat when angle_diff has a strict local minimum there whose drop from either neighbour is >= thres).
 * mean_dir_vec / angle_pairs ([n_pairs][2] joint indices) / change_angle are HOST arrays (the script's constants).
 * Stateless like ls_sag_decode; errors through ls_last_error(NULL).                                          */
#define LS_METRIC_MAX_JOINTS 64
#define LS_METRIC_MAX_PAIRS 8
int ls_motion_beats(int32_t B, int32_t njoints, int32_t n_frames, const float* sample,
                    const float* mean_dir_vec_host, const int32_t* angle_pairs_host,
                    const float* change_angle_host, int32_t n_pairs, float thres, float* angle_diff,
                    uint8_t* beat_mask, int32_t device, void* stream);
/* scripts/test_RAG_ted.py:112-123: per clip, sum over its audio onsets (audio_beats [B,M] seconds, n_audio [B], device)
 * of exp(-min_t (onset - t/fps)^2 / (2 sigma^2)) over the motion beats t of beat_mask, in fp64 like the script's numpy
 * arithmetic; clips without a motion beat score 0 and count no audio onsets (the script's `continue`).  Outputs
 * clip_score [B] fp64, clip_n_motion [B], clip_n_audio [B] (device); the caller sums them into
 * beat_align_score_sum / motion_beats_sum / num_beats.  The onset detector itself (librosa) stays with the caller.  */
int ls_beat_align(int32_t B, int32_t n_frames, const uint8_t* beat_mask, const float* audio_beats,
                  const int32_t* n_audio, int32_t M, float fps, float sigma, double* clip_score,
                  int32_t* clip_n_motion, int32_t* clip_n_audio, int32_t device, void* stream);

/* Loss terms of GaussianDiffusion.training_losses, HUBER branch (scripts/diffusion/gaussian_diffusion.py:21-24,
 * 1379-1391): target / output are dense [rows][n_frames] (rows = B * njoints * nfeats), z_mu / z_logvar n_z elements
 * (may both be NULL).  terms (device) receives {rot_mse, vel_mse, kld}: compute_huber of the samples, compute_huber of
 * their frame differences, -0.5 * mean(1 + logvar - mu^2 - exp(logvar)).  Stateless; errors through ls_last_error(NULL). */
int ls_huber_terms(int64_t rows, int32_t n_frames, const float* target, const float* output,
                   int64_t n_z, const float* z_mu, const float* z_logvar, float* terms,
                   int32_t device, void* stream);

/* Variational-bound terms of GaussianDiffusion._vb_terms_bpd / _prior_bpd (scripts/diffusion/gaussian_diffusion.py:1213-1247,
 * 1573-1590; normal_kl and discretized_gaussian_log_likelihood of scripts/diffusion/losses.py:12-77).  All tensors dense
 * [B][n] fp32 on the device, logvar1 / logvar2 one value per clip (LivelySpeaker's variances are fixed).  out[b] (device) =
 * mean_i normal_kl(mean1, logvar1, mean2, logvar2) / ln 2, or - where t is given and t[b] == 0 - the mean discretised
 * Gaussian negative log-likelihood of x_start under N(mean2, exp(logvar2)) / ln 2.  mean2 == NULL means 0 and logvar2 ==
 * NULL means 0 (the prior term; then t must be NULL).  Stateless; errors through ls_last_error(NULL).                   */
int ls_vb_terms(int32_t B, int64_t n, const float* x_start, const float* mean1, const float* mean2,
                const float* logvar1, const float* logvar2, const int64_t* t, float* out, int32_t device,
                void* stream);

/* ---- latent features of the evaluation's pose autoencoder (FGD features) ----------------
 * scripts/model/embedding_net.py:40-79 (PoseEncoderConv.forward in eval mode) as called by
 * scripts/model/ted_evaluator.py:35-41 (EmbeddingSpaceEvaluator.push_samples).  All pointers are DEVICE fp32.
 * Convolution weights keep torch's [out][in][k] layout; the linear weights are TRANSPOSED ([in][out]); every
 * BatchNorm (running statistics) is folded by the host to y = scale * x + shift per channel, applied after the bias.
 * slope_conv / slope_fc are the negative slopes of the LeakyReLUs of the two stacks (0.2 and - because the reference
 * writes nn.LeakyReLU(True) - 1.0).                                                                            */
typedef struct ls_pose_encoder_weights {
  int32_t pose_dim, n_frames;                 /* n_frames must be 34: the first linear layer is 384 = 32 x 12 wide */
  float slope_conv, slope_fc;
  const float *c1_w, *c1_b, *c1_scale, *c1_shift;   /* net.0: Conv1d(pose_dim, 32, 3) + BatchNorm1d(32)          */
  const float *c2_w, *c2_b, *c2_scale, *c2_shift;   /* net.1: Conv1d(32, 64, 3) + BatchNorm1d(64)                */
  const float *c3_w, *c3_b, *c3_scale, *c3_shift;   /* net.2: Conv1d(64, 64, 4, stride 2) + BatchNorm1d(64)      */
  const float *c4_w, *c4_b;                         /* net.3: Conv1d(64, 32, 3)                                  */
  const float *f1_wt, *f1_b, *f1_scale, *f1_shift;  /* out_net.0/1: Linear(384, 256) + BatchNorm1d(256)          */
  const float *f2_wt, *f2_b, *f2_scale, *f2_shift;  /* out_net.3/4: Linear(256, 128) + BatchNorm1d(128)          */
  const float *f3_wt, *f3_b;                        /* out_net.6: Linear(128, 32)                                */
  const float *mu_wt, *mu_b, *lv_wt, *lv_b;         /* fc_mu, fc_logvar: Linear(32, 32)                          */
} ls_pose_encoder_weights;
/* poses [B][n_frames][pose_dim] (device, the layout push_samples receives) -> mu [B][32] and, when not NULL,
 * logvar [B][32] (device).  Stateless; errors through ls_last_error(NULL).                                     */
int ls_pose_features(const ls_pose_encoder_weights* w, int32_t B, const float* poses, float* mu, float* logvar,
                     int32_t device, void* stream);

/* Introspection used by tests / bench: number of kernels launched by this handle
 * since creation, and read-back of the step-invariant buffers.                    */
int64_t ls_launch_count(const ls_handle* h);
/* which: 0 = A (audio proj) [B,34,512], 1 = P (prefix proj) [B,34,512],
 *        2 = z_mu [B,512], 3 = z_logvar [B,512], 4 = emb table [max_timestep,512] */
/* Copies min(capacity, n) floats into dst (device memory) and stores n in *n_elems.    */
int ls_debug_buffer(ls_handle* h, int32_t which, float* dst, int64_t capacity, int64_t* n_elems,
                    void* stream);
/* Per-layer parity aid for MLPblock / LN_spatial (scripts/model/mlp_module.py:29-35, 67-74): while dst != NULL, every
 * denoiser launch of this handle (ls_step, ls_step_multi's first step, ls_cfg_forward, ls_model_forward) also stores the
 * residual stream hidden[b][pass][token][512] (pass 0 = cond, 1 = uncond; S tokens) as it is AFTER MLPblock `layer`
 * (0-based), or right after the input projection / prefix tokens for layer = -1.  dst (device memory,
 * B*2*S*512 floats) = NULL switches it off.                                                                        */
int ls_debug_hidden(ls_handle* h, int32_t layer, float* dst);

#ifdef __cplusplus
}
#endif
#endif /* LIVELYSPEAKER_B200_H */
