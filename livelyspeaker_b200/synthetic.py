"""Seeded synthetic weights and conditioning for the RAG sampling path.

There are no checkpoints or datasets offline, so tests and bench.py use a
deterministic state_dict with the reference's exact key set / shapes
(scripts/model/RAG.py:17-74, scripts/model/mlp_module.py:37-91,
scripts/model/audio_enc.py:9-20) and synthetic conditioning of TED / BEAT shape
(SURVEY.md section 8d).

The reference's own init is degenerate for parity testing: channel-mix weights are
xavier * 1e-8 (mlp_module.py:63-65) and speaker / emotion embeddings are 1e-6
(RAG.py:67, scripts_beat/model/RAG.py:73), so a broken GEMM would still "pass".
This init therefore uses ordinary fan-in scaling everywhere and perturbs the LN
affine parameters away from (1, 0).
"""
import math
from dataclasses import dataclass

import torch

N_FRAMES = 34          # fixed by Conv1d(seq_len, seq_len, 1) and the WavEncoder stride chain
AUDIO_FEAT = 256       # WavEncoder output channels (audio_enc.py:18)
SPEAKER_DIM = 256      # RAG.py:66
N_SPEAKERS = 1400      # RAG.py:65
N_EMOTIONS = 8         # scripts_beat/model/RAG.py:72
PE_MAX_LEN = 5000      # mlp_module.py:105


@dataclass(frozen=True)
class RagDims:
    dataset: str
    njoints: int
    nfeats: int
    n_pre_emb: int      # style token (+ emotion token for BEAT)
    audio_len: int
    latent_dim: int = 512
    layers: int = 8

    @property
    def jd(self):
        return self.njoints * self.nfeats

    @property
    def seq_len(self):
        return N_FRAMES + self.n_pre_emb

    @property
    def in_feats(self):
        return 2 * self.jd + 1 + AUDIO_FEAT


TED = RagDims("ted", 9, 3, 1, 36267)
BEAT = RagDims("beat", 47, 6, 2, 36266)


def dims_for(dataset):
    return {"ted": TED, "beat": BEAT}[dataset]


def positional_table(d_model, max_len=PE_MAX_LEN):
    """The sinusoidal `pe` buffer, [max_len, 1, d] (mlp_module.py:105-113)."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(1).contiguous()


def synth_state_dict(dims=TED, seed=1):
    """Deterministic fp32 state_dict with the reference key set."""
    g = torch.Generator().manual_seed(seed)
    d, S = dims.latent_dim, dims.seq_len

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    def nrm(shape, std):
        return torch.randn(shape, generator=g) * std

    sd = {}
    for l in range(dims.layers):
        p = "backbone.mlps.%d." % l
        sd[p + "block1.0.alpha"] = 1 + nrm((1, 1, d), 0.1)
        sd[p + "block1.0.beta"] = nrm((1, 1, d), 0.1)
        sd[p + "block1.1.weight"] = uni((S, S, 1), S)
        sd[p + "block1.1.bias"] = uni((S,), S)
        sd[p + "block2.0.alpha"] = 1 + nrm((1, 1, d), 0.1)
        sd[p + "block2.0.beta"] = nrm((1, 1, d), 0.1)
        sd[p + "block2.1.weight"] = uni((d, d), d) * math.sqrt(3.0)   # ~xavier gain 1
        sd[p + "block2.1.bias"] = uni((d,), d)
    pe = positional_table(d)
    sd["backbone.sequence_pos_encoder.pe"] = pe
    sd["backbone.embed_timestep.sequence_pos_encoder.pe"] = pe
    for i in (0, 2):
        sd["backbone.embed_timestep.time_embed.%d.weight" % i] = uni((d, d), d)
        sd["backbone.embed_timestep.time_embed.%d.bias" % i] = uni((d,), d)
    sd["input_mapping.weight"] = uni((d, dims.in_feats), dims.in_feats)
    sd["input_mapping.bias"] = uni((d,), dims.in_feats)
    sd["sequence_pos_encoder.pe"] = pe.clone()
    sd["speaker_embedding.weight"] = nrm((N_SPEAKERS, SPEAKER_DIM), 0.1)
    for n in ("speaker_mu", "speaker_logvar"):
        sd[n + ".weight"] = uni((d, SPEAKER_DIM), SPEAKER_DIM)
        sd[n + ".bias"] = uni((d,), SPEAKER_DIM)
    if dims.n_pre_emb == 2:
        sd["emotion_embedding.weight"] = nrm((N_EMOTIONS, d), 0.1)
    for idx, (co, ci) in zip((0, 3, 6, 9), ((32, 1), (64, 32), (128, 64), (256, 128))):
        p = "audio_encoder.feat_extractor.%d." % idx
        sd[p + "weight"] = uni((co, ci, 15), ci * 15)
        sd[p + "bias"] = uni((co,), ci * 15)
    sd["output_process.poseFinal.weight"] = uni((dims.jd, d), d)
    sd["output_process.poseFinal.bias"] = uni((dims.jd,), d)
    return sd


def synth_sag_state_dict(seed=3, njoints=9, nfeats=3, d=512, ff=1024, layers=3):
    """Deterministic fp32 state_dict with the key set of the reference's SAG decoder
    (scripts/model/motionclip_module.py:98-134: nn.TransformerDecoder, 3 layers, + mapping / finallayer / pe)."""
    g = torch.Generator().manual_seed(seed)

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    def nrm(shape, std):
        return torch.randn(shape, generator=g) * std

    jd = njoints * nfeats
    sd = {"sequence_pos_encoder.pe": positional_table(d)}
    for l in range(layers):
        p = "seqTransDecoder.layers.%d." % l
        for a in ("self_attn.", "multihead_attn."):
            sd[p + a + "in_proj_weight"] = uni((3 * d, d), d) * math.sqrt(3.0)
            sd[p + a + "in_proj_bias"] = nrm((3 * d,), 0.02)
            sd[p + a + "out_proj.weight"] = uni((d, d), d)
            sd[p + a + "out_proj.bias"] = nrm((d,), 0.02)
        sd[p + "linear1.weight"] = uni((ff, d), d)
        sd[p + "linear1.bias"] = uni((ff,), d)
        sd[p + "linear2.weight"] = uni((d, ff), ff)
        sd[p + "linear2.bias"] = uni((d,), ff)
        for n in ("norm1", "norm2", "norm3"):
            sd[p + n + ".weight"] = 1 + nrm((d,), 0.1)
            sd[p + n + ".bias"] = nrm((d,), 0.1)
    sd["finallayer.weight"] = uni((jd, d), d)
    sd["finallayer.bias"] = uni((jd,), d)
    sd["mapping.weight"] = uni((d, jd + 1), jd + 1)
    sd["mapping.bias"] = uni((d,), jd + 1)
    return sd


def synth_embed_state_dict(seed=5, pose_dim=27):
    """Deterministic state_dict with the key set of the evaluation's EmbeddingNet(pose_dim, 34)
    (scripts/model/embedding_net.py:40-64, 166-216, 266-270 - the `gen_dict` of the autoencoder checkpoint
    scripts/model/ted_evaluator.py:16-20 loads), BatchNorm running statistics away from their defaults."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    def layer(name, shape):
        fan_in = 1
        for n in shape[1:]:
            fan_in *= n
        sd[name + ".weight"] = uni(shape, fan_in) * math.sqrt(3.0)
        sd[name + ".bias"] = uni((shape[0],), fan_in)

    def bn(name, c):
        sd[name + ".weight"] = 1 + 0.2 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.int64)

    e = "pose_encoder."
    for i, shape in enumerate(((32, pose_dim, 3), (64, 32, 3), (64, 64, 4))):
        layer(e + "net.%d.0" % i, shape)
        bn(e + "net.%d.1" % i, shape[0])
    layer(e + "net.3", (32, 64, 3))
    layer(e + "out_net.0", (256, 384))
    bn(e + "out_net.1", 256)
    layer(e + "out_net.3", (128, 256))
    bn(e + "out_net.4", 128)
    layer(e + "out_net.6", (32, 128))
    layer(e + "fc_mu", (32, 32))
    layer(e + "fc_logvar", (32, 32))
    d = "decoder."
    layer(d + "pre_net.0", (64, 32))
    bn(d + "pre_net.1", 64)
    layer(d + "pre_net.3", (136, 64))
    sd[d + "net.0.weight"] = uni((4, 32, 3), 12)           # ConvTranspose1d: [in, out, k]
    sd[d + "net.0.bias"] = uni((32,), 12)
    bn(d + "net.1", 32)
    sd[d + "net.3.weight"] = uni((32, 32, 3), 96)
    sd[d + "net.3.bias"] = uni((32,), 96)
    bn(d + "net.4", 32)
    layer(d + "net.6", (32, 32, 3))
    layer(d + "net.7", (pose_dim, 32, 3))
    return sd


def synth_cond(dims, batch, seed=233, scale=1.5, device="cpu"):
    """The `y` dict the eval scripts build (scripts/test_RAG_ted.py:63-70,
    scripts_beat/test_RAG_beat.py:100-125), synthetic values."""
    g = torch.Generator().manual_seed(seed)
    y = {
        "audio_input": 0.1 * torch.randn(batch, dims.audio_len, generator=g),
        "origin_x": 0.3 * torch.randn(batch, dims.njoints, dims.nfeats, N_FRAMES, generator=g),
        "scale": torch.full((batch,), float(scale)),
    }
    if dims.dataset == "ted":
        y["vid_indices"] = torch.randint(0, 1370, (batch,), generator=g)
    else:
        ids = torch.tensor([2, 4, 6, 8])
        y["vid_indices"] = ids[torch.randint(0, 4, (batch,), generator=g)]
        y["emo"] = torch.randint(0, N_EMOTIONS, (batch, 1), generator=g).repeat(1, N_FRAMES)
    y["mask"] = torch.ones(batch, N_FRAMES, dtype=torch.bool)
    y["lengths"] = torch.full((batch, N_FRAMES), float(N_FRAMES))
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in y.items()}
