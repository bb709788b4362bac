// C-ABI entry points of libls_b200.so (declared in include/livelyspeaker_b200.h).
#include <algorithm>
#include <cstdarg>
#include <cstring>

#include <cstddef>

#include "ls_internal.cuh"

static std::string g_create_error;

int ls_fail(ls_handle* h, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h)
    h->err = buf;
  else
    g_create_error = buf;
  return code;
}

static int dev_alloc(ls_handle* h, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) return ls_fail(h, LS_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  h->allocs.push_back(*p);
  return LS_OK;
}
template <class T>
static int dev_alloc_t(ls_handle* h, T** p, size_t n) { return dev_alloc(h, reinterpret_cast<void**>(p), n * sizeof(T)); }

static void add_raw(ls_handle* h, const std::string& key, std::vector<int64_t> shape, bool required = true) {
  RawTensor r;
  r.key = key;
  r.shape = std::move(shape);
  r.numel = 1;
  for (auto d : r.shape) r.numel *= d;
  r.required = required;
  h->raw.push_back(std::move(r));
}

static RawTensor* find_raw(ls_handle* h, const char* key) {
  for (auto& r : h->raw)
    if (r.key == key) return &r;
  return nullptr;
}

extern "C" int ls_abi_version(void) { return LS_ABI_VERSION; }

extern "C" const char* ls_last_error(const ls_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int ls_create(ls_handle** out, const ls_config* cfg) {
  if (!out || !cfg) return ls_fail(nullptr, LS_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->latent_dim != LS_D) return ls_fail(nullptr, LS_EUNSUPPORTED, "latent_dim %d (built for %d)", cfg->latent_dim, LS_D);
  if (cfg->n_frames != LS_F) return ls_fail(nullptr, LS_EUNSUPPORTED, "n_frames %d (the model fixes %d)", cfg->n_frames, LS_F);
  if (cfg->n_pre_emb != 1 && cfg->n_pre_emb != 2) return ls_fail(nullptr, LS_EINVAL, "n_pre_emb must be 1 or 2");
  if (cfg->n_layers < 1 || cfg->n_layers > LS_MAX_LAYERS) return ls_fail(nullptr, LS_EINVAL, "n_layers out of range");
  if (cfg->njoints < 1 || cfg->nfeats < 1 || cfg->max_batch < 1 || cfg->max_timestep < 1 || cfg->max_timestep > 5000)
    return ls_fail(nullptr, LS_EINVAL, "bad geometry");
  if (cfg->n_pre_emb == 2 && cfg->n_emotions < 1) return ls_fail(nullptr, LS_EINVAL, "n_emotions required for BEAT");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev)
    return ls_fail(nullptr, LS_ECUDA, "CUDA device %d not available (count %d)", cfg->device, ndev);
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, cfg->device);
  if (prop.major != 10)
    return ls_fail(nullptr, LS_EUNSUPPORTED, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name,
                   prop.major, prop.minor);
  if (cudaSetDevice(cfg->device) != cudaSuccess) return ls_fail(nullptr, LS_ECUDA, "cudaSetDevice failed");

  ls_handle* h = new ls_handle();
  h->cfg = *cfg;
  h->JD = cfg->njoints * cfg->nfeats;
  h->S = LS_F + cfg->n_pre_emb;
  const int64_t D = LS_D, S = h->S, JD = h->JD;
  for (int l = 0; l < cfg->n_layers; ++l) {
    std::string p = "backbone.mlps." + std::to_string(l) + ".";
    add_raw(h, p + "block1.0.alpha", {1, 1, D});
    add_raw(h, p + "block1.0.beta", {1, 1, D});
    add_raw(h, p + "block1.1.weight", {S, S, 1});
    add_raw(h, p + "block1.1.bias", {S});
    add_raw(h, p + "block2.0.alpha", {1, 1, D});
    add_raw(h, p + "block2.0.beta", {1, 1, D});
    add_raw(h, p + "block2.1.weight", {D, D});
    add_raw(h, p + "block2.1.bias", {D});
  }
  add_raw(h, "backbone.embed_timestep.sequence_pos_encoder.pe", {cfg->max_timestep, 1, D});
  add_raw(h, "backbone.embed_timestep.time_embed.0.weight", {D, D});
  add_raw(h, "backbone.embed_timestep.time_embed.0.bias", {D});
  add_raw(h, "backbone.embed_timestep.time_embed.2.weight", {D, D});
  add_raw(h, "backbone.embed_timestep.time_embed.2.bias", {D});
  add_raw(h, "input_mapping.weight", {D, 2 * JD + 1 + LS_AF});
  add_raw(h, "input_mapping.bias", {D});
  add_raw(h, "speaker_embedding.weight", {cfg->n_speakers, LS_SPK});
  add_raw(h, "speaker_mu.weight", {D, LS_SPK});
  add_raw(h, "speaker_mu.bias", {D});
  add_raw(h, "speaker_logvar.weight", {D, LS_SPK});
  add_raw(h, "speaker_logvar.bias", {D});
  if (cfg->n_pre_emb == 2) add_raw(h, "emotion_embedding.weight", {cfg->n_emotions, D});
  const int64_t conv[4][3] = {{0, 32, 1}, {3, 64, 32}, {6, 128, 64}, {9, 256, 128}};
  for (auto& c : conv) {
    std::string p = "audio_encoder.feat_extractor." + std::to_string(c[0]) + ".";
    add_raw(h, p + "weight", {c[1], c[2], 15});
    add_raw(h, p + "bias", {c[1]});
  }
  add_raw(h, "output_process.poseFinal.weight", {JD, D});
  add_raw(h, "output_process.poseFinal.bias", {JD});

  int rc = LS_OK;
  for (auto& r : h->raw)
    if ((rc = dev_alloc_t(h, &r.dev, (size_t)r.numel)) != LS_OK) break;
  const size_t MB = cfg->max_batch;
  h->wav_chunk = (int)std::min<size_t>(MB, 512);   // clips per WavEncoder pass: 0.7 GB of workspace at 512, grids that fill the GPU
  const size_t L1 = (cfg->audio_len + 3200 - 15) / 5 + 1, L2 = (L1 - 15) / 6 + 1, L3 = (L2 - 15) / 6 + 1;
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->A, MB * LS_F * LS_D);
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->P, MB * LS_F * LS_D);
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->z_mu, MB * LS_D);
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->z_lv, MB * LS_D);
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->emo_tok, MB * LS_D);
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->out_c, MB * JD * LS_F);
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->out_u, MB * JD * LS_F);
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->wav_a, (size_t)h->wav_chunk * std::max(32 * L1, 128 * L3));
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->wav_b, (size_t)h->wav_chunk * 64 * L2);
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->af, (size_t)h->wav_chunk * LS_AF * LS_F);
  if (rc == LS_OK) rc = dev_alloc_t(h, &h->t_tmp, MB);
  if (rc != LS_OK) {
    g_create_error = h->err;
    ls_destroy(h);
    return rc;
  }
  *out = h;
  return LS_OK;
}

extern "C" void ls_destroy(ls_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  lsf_destroy(h);
  lsw_destroy(h);
  for (void* p : {(void*)h->grad_ckpt, (void*)h->grad_gx, (void*)h->grad_t})
    if (p) cudaFree(p);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

static bool ignorable_key(const char* key) {
  return strncmp(key, "clip_model.", 11) == 0 || strcmp(key, "backbone.sequence_pos_encoder.pe") == 0 ||
         strcmp(key, "sequence_pos_encoder.pe") == 0;
}

extern "C" int ls_load_weight(ls_handle* h, const char* key, const float* ptr, const int64_t* shape, int32_t ndim,
                              void* stream) {
  if (!h || !key || !ptr || !shape) return ls_fail(h, LS_EINVAL, "null argument");
  if (ignorable_key(key)) return LS_OK;
  RawTensor* r = find_raw(h, key);
  if (!r) return ls_fail(h, LS_EINVAL, "unexpected key '%s'", key);
  bool ok = (size_t)ndim == r->shape.size();
  for (int i = 0; ok && i < ndim; ++i) {
    // the pe buffer may be longer than the table we use (reference: 5000 rows)
    if (i == 0 && r->key == "backbone.embed_timestep.sequence_pos_encoder.pe")
      ok = shape[0] >= r->shape[0];
    else
      ok = shape[i] == r->shape[i];
  }
  if (!ok) return ls_fail(h, LS_EINVAL, "shape mismatch for '%s'", key);
  LS_CUDA(h, cudaSetDevice(h->cfg.device));
  LS_CUDA(h, cudaMemcpyAsync(r->dev, ptr, (size_t)r->numel * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  r->loaded = true;
  h->finalized = false;
  return LS_OK;
}

extern "C" int ls_finalize_weights(ls_handle* h, void* stream) {
  if (!h) return LS_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  LS_CUDA(h, cudaSetDevice(h->cfg.device));
  for (auto& r : h->raw)
    if (r.required && !r.loaded) return ls_fail(h, LS_ESTATE, "missing key '%s'", r.key.c_str());
  auto R = [&](const std::string& k) { return find_raw(h, k.c_str())->dev; };
  const int D = LS_D, JD = h->JD, IN = 2 * JD + 1 + LS_AF;
  int rc;
  // derived buffers are allocated once and rewritten on every finalize
  if (h->w.emb_table == nullptr) {
    float* p;
    for (int l = 0; l < h->cfg.n_layers; ++l) {
      if ((rc = dev_alloc_t(h, &p, (size_t)D * D))) return rc;
      h->w.layer[l].w_ch_t = p;
    }
    if ((rc = dev_alloc_t(h, &p, (size_t)JD * D))) return rc;
    h->w.w_x_t = p;
    if ((rc = dev_alloc_t(h, &p, (size_t)JD * D))) return rc;
    h->w.w_o_t = p;
    if ((rc = dev_alloc_t(h, &p, (size_t)D))) return rc;
    h->w.w_bit = p;
    if ((rc = dev_alloc_t(h, &p, (size_t)LS_AF * D))) return rc;
    h->w.w_a_t = p;
    if ((rc = dev_alloc_t(h, &p, (size_t)LS_SPK * D))) return rc;
    h->w.w_mu_t = p;
    if ((rc = dev_alloc_t(h, &p, (size_t)LS_SPK * D))) return rc;
    h->w.w_lv_t = p;
    if ((rc = dev_alloc_t(h, &p, (size_t)h->cfg.max_timestep * D))) return rc;
    h->w.emb_table = p;
    if ((rc = dev_alloc_t(h, &h->w1_t, (size_t)D * D))) return rc;
    if ((rc = dev_alloc_t(h, &h->w2_t, (size_t)D * D))) return rc;
  }
  float *w1t = h->w1_t, *w2t = h->w2_t;   // transposed time-embedding MLP weights

  for (int l = 0; l < h->cfg.n_layers; ++l) {
    std::string p = "backbone.mlps." + std::to_string(l) + ".";
    LsLayerW& L = h->w.layer[l];
    L.ln1_a = R(p + "block1.0.alpha");
    L.ln1_b = R(p + "block1.0.beta");
    L.w_tok = R(p + "block1.1.weight");
    L.b_tok = R(p + "block1.1.bias");
    L.ln2_a = R(p + "block2.0.alpha");
    L.ln2_b = R(p + "block2.0.beta");
    L.b_ch = R(p + "block2.1.bias");
    if ((rc = lsk_transpose(h, R(p + "block2.1.weight"), const_cast<float*>(L.w_ch_t), D, D, D, s))) return rc;
  }
  const float* win = R("input_mapping.weight");
  if ((rc = lsk_transpose(h, win, const_cast<float*>(h->w.w_x_t), D, JD, IN, s))) return rc;
  if ((rc = lsk_transpose(h, win + JD, const_cast<float*>(h->w.w_o_t), D, JD, IN, s))) return rc;
  if ((rc = lsk_transpose(h, win + 2 * JD, const_cast<float*>(h->w.w_bit), D, 1, IN, s))) return rc;
  if ((rc = lsk_transpose(h, win + 2 * JD + 1, const_cast<float*>(h->w.w_a_t), D, LS_AF, IN, s))) return rc;
  h->w.b_in = R("input_mapping.bias");
  h->w.w_out = R("output_process.poseFinal.weight");
  h->w.b_out = R("output_process.poseFinal.bias");
  h->w.spk_emb = R("speaker_embedding.weight");
  if ((rc = lsk_transpose(h, R("speaker_mu.weight"), const_cast<float*>(h->w.w_mu_t), D, LS_SPK, LS_SPK, s))) return rc;
  if ((rc = lsk_transpose(h, R("speaker_logvar.weight"), const_cast<float*>(h->w.w_lv_t), D, LS_SPK, LS_SPK, s))) return rc;
  h->w.b_mu = R("speaker_mu.bias");
  h->w.b_lv = R("speaker_logvar.bias");
  h->w.emo_emb = h->cfg.n_pre_emb == 2 ? R("emotion_embedding.weight") : nullptr;
  const std::string te = "backbone.embed_timestep.";
  if ((rc = lsk_transpose(h, R(te + "time_embed.0.weight"), w1t, D, D, D, s))) return rc;
  if ((rc = lsk_transpose(h, R(te + "time_embed.2.weight"), w2t, D, D, D, s))) return rc;
  if ((rc = lsk_time_embed_table(h, R(te + "sequence_pos_encoder.pe"), w1t, R(te + "time_embed.0.bias"), w2t,
                                 R(te + "time_embed.2.bias"), const_cast<float*>(h->w.emb_table), h->cfg.max_timestep, s)))
    return rc;
  if ((rc = lsf_init(h, s)) < 0) return rc;
  if (rc == 0) {      // tcgen05 path available: the WavEncoder's three wide convolutions run on the tensor cores too
    const float* wc[3] = {R("audio_encoder.feat_extractor.3.weight"), R("audio_encoder.feat_extractor.6.weight"),
                          R("audio_encoder.feat_extractor.9.weight")};
    if ((rc = lsw_init(h, wc, s, R("audio_encoder.feat_extractor.0.weight"))) < 0) return rc;
  }
  h->finalized = true;
  h->cond_batch = 0;
  return LS_OK;
}

extern "C" int ls_set_impl(ls_handle* h, int32_t impl) {
  if (!h) return LS_EINVAL;
  if (impl < LS_IMPL_AUTO || impl > LS_IMPL_TC_BF16) return ls_fail(h, LS_EINVAL, "unknown impl %d", impl);
  if ((impl == LS_IMPL_TC_BF16X3 || impl == LS_IMPL_TC_BF16) && !lsf_available(h))
    return ls_fail(h, LS_EUNSUPPORTED, "the tcgen05 path is not available in this build / for this geometry");
  h->impl = impl;
  return LS_OK;
}

extern "C" int ls_get_impl(const ls_handle* h) {
  if (!h) return LS_EINVAL;
  if (h->impl != LS_IMPL_AUTO) return h->impl;
  return lsf_available(h) ? LS_IMPL_TC_BF16X3 : LS_IMPL_SIMT;
}

static int check_ready(ls_handle* h, int B, bool need_cond) {
  if (!h) return LS_EINVAL;
  if (!h->finalized) return ls_fail(h, LS_ESTATE, "weights not finalized");
  if (B < 1 || B > h->cfg.max_batch) return ls_fail(h, LS_EINVAL, "batch %d outside [1,%d]", B, h->cfg.max_batch);
  if (need_cond && h->cond_batch != B)
    return ls_fail(h, LS_ESTATE, "ls_precompute_cond was run for batch %d, this call has %d", h->cond_batch, B);
  cudaError_t e = cudaSetDevice(h->cfg.device);
  if (e != cudaSuccess) return ls_fail(h, LS_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return LS_OK;
}

extern "C" int ls_wav_encoder(ls_handle* h, int32_t B, const float* audio, float* out, void* stream) {
  int rc = check_ready(h, B, false);
  if (rc) return rc;
  if (!audio || !out) return ls_fail(h, LS_EINVAL, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  for (int c0 = 0; c0 < B; c0 += h->wav_chunk) {
    const int nb = std::min(h->wav_chunk, B - c0);
    if ((rc = lsk_wav_encoder(h, nb, audio + (size_t)c0 * h->cfg.audio_len, h->af, s))) return rc;
    if ((rc = lsk_cm_to_fm(h, h->af, out + (size_t)c0 * LS_F * LS_AF, nb, s))) return rc;
  }
  return LS_OK;
}

extern "C" int ls_precompute_cond(ls_handle* h, int32_t B, const float* audio, float* origin_x,
                                  const int64_t* vid_indices, const int64_t* emo, int64_t emo_stride,
                                  int32_t mutate_origin, void* stream) {
  int rc = check_ready(h, B, false);
  if (rc) return rc;
  if (!audio || !origin_x || !vid_indices) return ls_fail(h, LS_EINVAL, "null argument");
  if (h->cfg.n_pre_emb == 2 && !emo) return ls_fail(h, LS_EINVAL, "BEAT geometry needs y['emo']");
  cudaStream_t s = (cudaStream_t)stream;
  for (int c0 = 0; c0 < B; c0 += h->wav_chunk) {
    const int nb = std::min(h->wav_chunk, B - c0);
    if ((rc = lsk_wav_encoder(h, nb, audio + (size_t)c0 * h->cfg.audio_len, h->af, s))) return rc;
    if ((rc = lsk_cond_proj(h, nb, c0, h->af, origin_x, vid_indices, emo, emo_stride, mutate_origin, s))) return rc;
  }
  h->cond_batch = B;
  return LS_OK;
}

// One pass of RAG.forward.  On the tensor-core implementations it is the fused kernel in mode 2: the kernel always runs
// both guidance passes of a clip (they share every weight block) and combines them as out_u + scale (out_c - out_u), so
// scale = 0 returns the uncond pass exactly and scale = 1 the cond pass to within an ulp - 2x the arithmetic of a
// single pass at 10x the speed of the fp32 CUDA-core kernel, which LS_IMPL_SIMT keeps as the exact-order path.
static int forward_one_pass(ls_handle* h, int B, const float* x, const int64_t* t, int sel, const uint8_t* cond_drop,
                            const float* style_eps, float* out, cudaStream_t s) {
  const int impl = ls_get_impl(h);
  if (impl == LS_IMPL_TC_BF16X3 || impl == LS_IMPL_TC_BF16) {
    float* scale = h->out_c;          // [B] of a scratch buffer the tensor-core path does not use
    int rc = lsk_pass_scale(h, B, sel == 1, cond_drop, scale, s);
    if (rc) return rc;
    ls_step_params p{};
    p.mode = 2;
    const ls_step_io io{style_eps, style_eps, nullptr, 0, 0, 0, nullptr, out};
    return lsf_steps(h, B, 1, &p, &io, impl == LS_IMPL_TC_BF16X3, x, scale, s, t);
  }
  return lsk_denoise_simt(h, B, x, t, -1, sel, style_eps, style_eps, out, out, s, nullptr, cond_drop);
}

extern "C" int ls_model_forward(ls_handle* h, int32_t B, const float* x, const int64_t* t, int32_t uncond,
                                const float* style_eps, float* out, float* z_mu, float* z_logvar, void* stream) {
  int rc = check_ready(h, B, true);
  if (rc) return rc;
  if (!x || !t || !style_eps || !out) return ls_fail(h, LS_EINVAL, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = forward_one_pass(h, B, x, t, uncond ? 2 : 1, nullptr, style_eps, out, s))) return rc;
  if (z_mu) LS_CUDA(h, cudaMemcpyAsync(z_mu, h->z_mu, (size_t)B * LS_D * 4, cudaMemcpyDeviceToDevice, s));
  if (z_logvar) LS_CUDA(h, cudaMemcpyAsync(z_logvar, h->z_lv, (size_t)B * LS_D * 4, cudaMemcpyDeviceToDevice, s));
  return LS_OK;
}

extern "C" int ls_model_forward_train(ls_handle* h, int32_t B, const float* x, const int64_t* t, const uint8_t* cond_drop,
                                      const float* style_eps, float* out, float* z_mu, float* z_logvar, void* stream) {
  int rc = check_ready(h, B, true);
  if (rc) return rc;
  if (!x || !t || !style_eps || !out) return ls_fail(h, LS_EINVAL, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = forward_one_pass(h, B, x, t, 1, cond_drop, style_eps, out, s))) return rc;
  if (z_mu) LS_CUDA(h, cudaMemcpyAsync(z_mu, h->z_mu, (size_t)B * LS_D * 4, cudaMemcpyDeviceToDevice, s));
  if (z_logvar) LS_CUDA(h, cudaMemcpyAsync(z_logvar, h->z_lv, (size_t)B * LS_D * 4, cudaMemcpyDeviceToDevice, s));
  return LS_OK;
}

extern "C" int ls_cfg_forward(ls_handle* h, int32_t B, const float* x, const int64_t* t, const float* eps_cond,
                              const float* eps_uncond, const float* scale, float* out, void* stream) {
  int rc = check_ready(h, B, true);
  if (rc) return rc;
  if (!x || !t || !eps_cond || !eps_uncond || !scale || !out) return ls_fail(h, LS_EINVAL, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int impl = ls_get_impl(h);
  if (impl == LS_IMPL_TC_BF16X3 || impl == LS_IMPL_TC_BF16) {
    // the fused tcgen05 kernel in mode 2: both passes + guidance, per-clip timesteps read on the device, no update
    ls_step_params p{};
    p.mode = 2;
    const ls_step_io io{eps_cond, eps_uncond, nullptr, 0, 0, 0, nullptr, out};
    return lsf_steps(h, B, 1, &p, &io, impl == LS_IMPL_TC_BF16X3, x, scale, s, t);
  }
  if ((rc = lsk_denoise_simt(h, B, x, t, -1, 3, eps_cond, eps_uncond, h->out_c, h->out_u, s))) return rc;
  return lsk_cfg_combine(h, B, h->out_c, h->out_u, scale, out, s);
}

// ---- differentiable denoiser call (the *_with_grad samplers) ------------------------------------------------------
static int grad_workspace(ls_handle* h, int B) {
  if (B <= h->grad_cap) return LS_OK;
  for (void* p : {(void*)h->grad_ckpt, (void*)h->grad_gx, (void*)h->grad_t})
    if (p) cudaFree(p);
  h->grad_ckpt = nullptr; h->grad_gx = nullptr; h->grad_t = nullptr; h->grad_cap = 0;
  const size_t ck = (size_t)B * 2 * h->cfg.n_layers * h->S * LS_D * sizeof(float);
  if (cudaMalloc(&h->grad_ckpt, ck) != cudaSuccess || cudaMalloc(&h->grad_gx, (size_t)2 * B * h->JD * LS_F * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&h->grad_t, (size_t)B * sizeof(int64_t)) != cudaSuccess)
    return ls_fail(h, LS_ENOMEM, "ls_cfg_forward_grad: %zu bytes of checkpoints for batch %d", ck, B);
  h->grad_cap = B;
  return LS_OK;
}

extern "C" int ls_cfg_forward_grad(ls_handle* h, int32_t B, const float* x, const int64_t* t, const float* eps_cond,
                                   const float* eps_uncond, const float* scale, float* out, void* stream) {
  int rc = check_ready(h, B, true);
  if (rc) return rc;
  if (!x || !t || !eps_cond || !eps_uncond || !scale || !out) return ls_fail(h, LS_EINVAL, "null argument");
  if ((rc = grad_workspace(h, B))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  LS_CUDA(h, cudaMemcpyAsync(h->grad_t, t, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  if ((rc = lsk_denoise_simt(h, B, x, t, -1, 3, eps_cond, eps_uncond, h->out_c, h->out_u, s, h->grad_ckpt))) return rc;
  h->grad_batch = B;
  return lsk_cfg_combine(h, B, h->out_c, h->out_u, scale, out, s);
}

extern "C" int ls_cfg_backward(ls_handle* h, int32_t B, const float* grad_out, const float* scale, float* grad_x,
                               void* stream) {
  int rc = check_ready(h, B, true);
  if (rc) return rc;
  if (!grad_out || !scale || !grad_x) return ls_fail(h, LS_EINVAL, "null argument");
  if (h->grad_batch != B) return ls_fail(h, LS_ESTATE, "ls_cfg_backward: no saved forward for batch %d (last: %d)", B, h->grad_batch);
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = lsk_denoise_simt_bwd(h, B, h->grad_t, h->grad_ckpt, grad_out, scale, h->grad_gx, s))) return rc;
  const int64_t n = (int64_t)B * h->JD * LS_F;
  return lsk_axpby(h, n, h->grad_gx, h->grad_gx + n, 1.f, 1.f, grad_x, s);
}

static_assert(sizeof(ls_step_params) == 48 && sizeof(ls_step_io) == 64 && offsetof(ls_step_io, x_prev) == 48,
              "ls_step_params / ls_step_io layouts are part of the ABI (livelyspeaker_b200/_cabi.py mirrors them)");

static int check_step(ls_handle* h, const ls_step_params* p, const ls_step_io* io, bool allow_mode2) {
  if (p->mode < 0 || p->mode > (allow_mode2 ? 2 : 1)) return ls_fail(h, LS_EINVAL, "mode %d", p->mode);
  if (!io->eps_cond || !io->eps_uncond) return ls_fail(h, LS_EINVAL, "null argument");
  if (p->mode != 2 && !io->x_prev) return ls_fail(h, LS_EINVAL, "x_prev is NULL");
  if (p->mode == 2 && !io->pred_x0) return ls_fail(h, LS_EINVAL, "mode 2 needs pred_x0");
  if (p->mode != 2 && p->add_noise && !io->noise) return ls_fail(h, LS_EINVAL, "noise is NULL");
  if (p->t_model < 0 || p->t_model >= h->cfg.max_timestep)
    return ls_fail(h, LS_EINVAL, "t_model %d outside the embedding table [0,%d)", p->t_model, h->cfg.max_timestep);
  return LS_OK;
}

static int step_simt(ls_handle* h, int B, const ls_step_params* p, const ls_step_io* io, const float* x_t,
                     const float* scale, cudaStream_t s) {
  int rc;
  if ((rc = lsk_denoise_simt(h, B, x_t, nullptr, p->t_model, 3, io->eps_cond, io->eps_uncond, h->out_c, h->out_u, s)))
    return rc;
  return lsk_cfg_update(h, B, p, h->out_c, h->out_u, scale, x_t, io->noise, io->noise_sb, io->noise_sj, io->noise_sf,
                        io->x_prev, io->pred_x0, s);
}

extern "C" int ls_step(ls_handle* h, int32_t B, const ls_step_params* p, const float* x_t, const float* eps_cond,
                       const float* eps_uncond, const float* noise, int64_t noise_sb, int64_t noise_sj,
                       int64_t noise_sf, const float* scale, float* x_prev, float* pred_x0, void* stream) {
  int rc = check_ready(h, B, true);
  if (rc) return rc;
  if (!p || !x_t || !scale) return ls_fail(h, LS_EINVAL, "null argument");
  const ls_step_io io{eps_cond, eps_uncond, noise, noise_sb, noise_sj, noise_sf, x_prev, pred_x0};
  if ((rc = check_step(h, p, &io, true))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int impl = ls_get_impl(h);
  if (impl == LS_IMPL_TC_BF16X3 || impl == LS_IMPL_TC_BF16)
    return lsf_steps(h, B, 1, p, &io, impl == LS_IMPL_TC_BF16X3, x_t, scale, s);
  return step_simt(h, B, p, &io, x_t, scale, s);
}

extern "C" int ls_step_multi(ls_handle* h, int32_t B, int32_t n_steps, const ls_step_params* p, const ls_step_io* io,
                             const float* x_t, const float* scale, void* stream) {
  int rc = check_ready(h, B, true);
  if (rc) return rc;
  if (!p || !io || !x_t || !scale) return ls_fail(h, LS_EINVAL, "null argument");
  if (n_steps < 1 || n_steps > LS_MAX_FUSED_STEPS)
    return ls_fail(h, LS_EINVAL, "n_steps %d outside [1,%d]", n_steps, LS_MAX_FUSED_STEPS);
  for (int k = 0; k < n_steps; ++k) {
    if ((rc = check_step(h, p + k, io + k, false))) return rc;
    if (n_steps > 1) {
      if (io[k].x_prev == x_t) return ls_fail(h, LS_EINVAL, "x_prev of step %d aliases x_t", k);
      for (int j = 0; j < k; ++j)
        if (io[j].x_prev == io[k].x_prev) return ls_fail(h, LS_EINVAL, "steps %d and %d share x_prev", j, k);
    }
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int impl = ls_get_impl(h);
  if (impl == LS_IMPL_TC_BF16X3 || impl == LS_IMPL_TC_BF16)
    return lsf_steps(h, B, n_steps, p, io, impl == LS_IMPL_TC_BF16X3, x_t, scale, s);
  for (int k = 0; k < n_steps; ++k)
    if ((rc = step_simt(h, B, p + k, io + k, k == 0 ? x_t : io[k - 1].x_prev, scale, s))) return rc;
  return LS_OK;
}

extern "C" int ls_q_sample(ls_handle* h, int64_t n, const float* x0, const float* noise, float c_x0, float c_noise,
                           float* out, void* stream) {
  if (!h || !x0 || !noise || !out || n < 0) return ls_fail(h, LS_EINVAL, "bad argument");
  LS_CUDA(h, cudaSetDevice(h->cfg.device));
  if (n == 0) return LS_OK;
  return lsk_axpby(h, n, x0, noise, c_x0, c_noise, out, (cudaStream_t)stream);
}

extern "C" int64_t ls_launch_count(const ls_handle* h) { return h ? h->launches : -1; }

extern "C" int ls_debug_hidden(ls_handle* h, int32_t layer, float* dst) {
  if (!h) return LS_EINVAL;
  if (dst && (layer < -1 || layer >= h->cfg.n_layers)) return ls_fail(h, LS_EINVAL, "layer %d outside [-1,%d)", layer, h->cfg.n_layers);
  h->dbg_h = dst;
  h->dbg_layer = dst ? layer : -2;
  return LS_OK;
}

extern "C" int ls_debug_buffer(ls_handle* h, int32_t which, float* dst, int64_t capacity, int64_t* n_elems,
                               void* stream) {
  if (!h || !n_elems) return LS_EINVAL;
  const int64_t B = h->cond_batch;
  const float* src = nullptr;
  int64_t n = 0;
  switch (which) {
    case 0: src = h->A; n = B * LS_F * LS_D; break;
    case 1: src = h->P; n = B * LS_F * LS_D; break;
    case 2: src = h->z_mu; n = B * LS_D; break;
    case 3: src = h->z_lv; n = B * LS_D; break;
    case 4: src = h->w.emb_table; n = (int64_t)h->cfg.max_timestep * LS_D; break;
    default: return ls_fail(h, LS_EINVAL, "unknown buffer %d", which);
  }
  *n_elems = n;
  if (dst && capacity > 0 && n > 0)
    LS_CUDA(h, cudaMemcpyAsync(dst, src, (size_t)std::min(n, capacity) * sizeof(float), cudaMemcpyDeviceToDevice,
                               (cudaStream_t)stream));
  return LS_OK;
}
