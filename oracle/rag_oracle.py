"""Oracle (test infrastructure): the RAG denoiser, classifier-free wrapper.

Plain torch-CPU fp32 restatement, functional (state_dict in, tensors out), with
every random draw passed in explicitly so the CUDA path can be fed the very
same numbers.  Written in the reference's literal op order (no hoisting) -
the hoisted algebra lives in the CUDA kernels and is what these functions test.

Reference:
  scripts/model/audio_enc.py:6-25       WavEncoder
  scripts/model/mlp_module.py:21-35     LN_spatial
  scripts/model/mlp_module.py:37-74     MLPblock
  scripts/model/mlp_module.py:76-91     TransMLP
  scripts/model/mlp_module.py:123-136   TimestepEmbedder
  scripts/model/RAG.py:98-133           RAG.forward  (+176-211 In/OutputProcess)
  scripts/model/cfg_sampler.py:24-31    ClassifierFreeSampleModel.forward
  scripts_beat/model/RAG.py:56,72-74,119-126   BEAT twin (emotion token, S=36)
"""
import torch
import torch.nn.functional as F

N_PRE_SEQ = 4  # RAG.py:70


def wav_encoder(sd, audio):
    """audio [B, L] -> [B, 34, 256].  audio_enc.py:9-25 (InstanceNorm1d default:
    no affine, eps 1e-5, biased variance; LeakyReLU slope 0.3)."""
    p = "audio_encoder.feat_extractor."
    x = audio.unsqueeze(1)
    x = F.conv1d(x, sd[p + "0.weight"], sd[p + "0.bias"], stride=5, padding=1600)
    x = F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.3)
    x = F.conv1d(x, sd[p + "3.weight"], sd[p + "3.bias"], stride=6)
    x = F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.3)
    x = F.conv1d(x, sd[p + "6.weight"], sd[p + "6.bias"], stride=6)
    x = F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.3)
    x = F.conv1d(x, sd[p + "9.weight"], sd[p + "9.bias"], stride=6)
    return x.transpose(1, 2)


def ln_spatial(x, alpha, beta, eps=1e-5):
    """mlp_module.py:29-35."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * alpha + beta


def mlp_block(sd, layer, x, emb):
    """mlp_module.py:67-74 with block1 = LN,Conv1d(S,S,1),SiLU and
    block2 = LN,Linear,SiLU (:51-60)."""
    p = "backbone.mlps.%d." % layer
    x = x + emb
    u = ln_spatial(x, sd[p + "block1.0.alpha"], sd[p + "block1.0.beta"])
    u = F.conv1d(u, sd[p + "block1.1.weight"], sd[p + "block1.1.bias"])
    x = x + F.silu(u)
    u = ln_spatial(x, sd[p + "block2.0.alpha"], sd[p + "block2.0.beta"])
    u = F.linear(u, sd[p + "block2.1.weight"], sd[p + "block2.1.bias"])
    return x + F.silu(u)


def timestep_embed(sd, t):
    """mlp_module.py:135-136: t int64 [B] (ORIGINAL timesteps) -> [B,1,512]."""
    p = "backbone.embed_timestep."
    e = sd[p + "sequence_pos_encoder.pe"][t]
    e = F.linear(e, sd[p + "time_embed.0.weight"], sd[p + "time_embed.0.bias"])
    return F.linear(F.silu(e), sd[p + "time_embed.2.weight"], sd[p + "time_embed.2.bias"])


def n_layers(sd):
    n = 0
    while ("backbone.mlps.%d.block2.1.weight" % n) in sd:
        n += 1
    return n


def trans_mlp(sd, x, t, trace=None):
    """mlp_module.py:85-91.  `trace` (a list) receives the hidden state before the first block and after every block."""
    emb = timestep_embed(sd, t)
    if trace is not None:
        trace.append(x)
    for layer in range(n_layers(sd)):
        x = mlp_block(sd, layer, x, emb)
        if trace is not None:
            trace.append(x)
    return x


def rag_forward(sd, x, t, y, style_eps, njoints, nfeats, trace=None):
    """RAG.py:98-133 (TED) / scripts_beat/model/RAG.py:101-137 (BEAT, selected by
    the presence of 'emotion_embedding.weight').  style_eps replaces the
    randn_like of reparameterize (RAG.py:10-13).  Mutates y['origin_x'] in place
    exactly like the reference (RAG.py:110)."""
    bs, nj, nf, nframes = x.shape
    af = wav_encoder(sd, y["audio_input"])
    audio_emb = torch.zeros_like(af) if y.get("uncond", False) else af
    if y.get("cond_drop") is not None and not y.get("uncond", False):
        # training-mode mask_cond (RAG.py:84-93): 1 = this clip's condition is replaced by zeros
        audio_emb = (af.flatten(1) * (1.0 - y["cond_drop"].view(bs, 1))).reshape(af.shape)
    ox = y["origin_x"]
    ox[..., N_PRE_SEQ:] = 0
    xi = torch.cat([x, ox], dim=1)
    xi = xi.permute(3, 0, 1, 2).reshape(nframes, bs, 2 * nj * nf)
    bit = xi.new_zeros(nframes, bs, 1)
    bit[:N_PRE_SEQ] = 1
    xi = torch.cat([xi, bit], dim=-1).permute(1, 0, 2)
    xi = torch.cat([xi, audio_emb], dim=-1)
    h = F.linear(xi, sd["input_mapping.weight"], sd["input_mapping.bias"])
    z = sd["speaker_embedding.weight"][y["vid_indices"]][:, None]
    z_mu = F.linear(z, sd["speaker_mu.weight"], sd["speaker_mu.bias"])
    z_lv = F.linear(z, sd["speaker_logvar.weight"], sd["speaker_logvar.bias"])
    style = z_mu + style_eps * torch.exp(0.5 * z_lv)
    toks = [style]
    if "emotion_embedding.weight" in sd:
        toks.append(sd["emotion_embedding.weight"][y["emo"][:, 0]][:, None])
    n_pre = len(toks)
    h = torch.cat(toks + [h], dim=1)
    h = trans_mlp(sd, h, t, trace)[:, n_pre:].permute(1, 0, 2)
    o = F.linear(h, sd["output_process.poseFinal.weight"], sd["output_process.poseFinal.bias"])
    o = o.reshape(nframes, bs, njoints, nfeats).permute(1, 2, 3, 0)
    return {"output": o, "z_mu": z_mu, "z_logvar": z_lv}


def cfg_forward(sd, x, t, y, eps_cond, eps_uncond, njoints, nfeats):
    """cfg_sampler.py:24-31.  The cond call runs (and draws) first."""
    out = rag_forward(sd, x, t, y, eps_cond, njoints, nfeats)["output"]
    yu = dict(y)
    yu["uncond"] = True
    out_u = rag_forward(sd, x, t, yu, eps_uncond, njoints, nfeats)["output"]
    return out_u + y["scale"].view(-1, 1, 1, 1) * (out - out_u)
