"""`RAG` - the Rhythm-Aware-Gesture denoiser behind the reference's Python surface.

Mirrors scripts/model/RAG.py:16-133 (TED) and scripts_beat/model/RAG.py:16-137 (BEAT):
same constructor arguments, same attributes, same ``forward(x, timesteps, y)`` contract
and - most importantly - a ``state_dict()`` with exactly the reference's 88 (TED) / 89
(BEAT) keys and shapes, so ``torch.load`` + ``load_model_wo_clip`` of an existing
checkpoint works unchanged.

Unlike the reference this module holds no layer objects: parameters hang on an
anonymous module tree that only reproduces the key names, and ``forward`` hands raw
device pointers to the CUDA library through the C ABI (``_cabi.Engine``).  There is
no PyTorch implementation of the math in this package.
"""
import math

import torch
import torch.nn as nn

from . import _cabi
from .synthetic import (AUDIO_FEAT, N_EMOTIONS, N_FRAMES, N_SPEAKERS, SPEAKER_DIM, RagDims, positional_table)


def _attach(root, dotted, tensor, buffer=False, share=None):
    """Register `tensor` under the dotted state_dict name, creating plain container
    modules for the intermediate path components."""
    parts = dotted.split(".")
    mod = root
    for k, name in enumerate(parts[:-1]):
        child = mod._modules.get(name)
        if child is None:
            child = share if (share is not None and k == len(parts) - 2) else nn.Module()
            mod.add_module(name, child)
        mod = child
    if share is not None:
        return
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor))


class RAG(nn.Module):
    def __init__(self, modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot,
                 latent_dim=256, ff_size=1024, num_layers=8, num_heads=4, dropout=0.1, ablation=None,
                 activation="gelu", legacy=False, data_rep='rot6d', clip_dim=512, arch='trans_enc',
                 mlpact='silu', n_pre_emb=1, audio_len=None, **kargs):
        super().__init__()
        # attribute surface of the reference module (RAG.py:22-57); ff_size / num_heads /
        # activation are accepted and stored but unused there as well.
        self.legacy, self.modeltype = legacy, modeltype
        self.njoints, self.nfeats, self.num_actions = njoints, nfeats, num_actions
        self.data_rep, self.pose_rep, self.glob, self.glob_rot = data_rep, pose_rep, glob, glob_rot
        self.translation = translation
        self.cond_mode = kargs.get('cond_mode', 'no_cond')
        self.latent_dim, self.ff_size, self.num_layers, self.num_heads = latent_dim, ff_size, num_layers, num_heads
        self.dropout, self.ablation, self.activation, self.clip_dim = dropout, ablation, activation, clip_dim
        self.action_emb = kargs.get('action_emb', None)
        self.input_feats = njoints * nfeats
        self.cond_mask_prob = kargs.get('cond_mask_prob', 0.)
        self.arch, self.mlpact = arch, mlpact
        self.gru_emb_dim = latent_dim if arch == 'gru' else 0
        self.n_pre_seq = 4
        if mlpact != 'silu':
            raise NotImplementedError("only mlpact='silu' (the shipped configuration) is built into the kernels")
        if n_pre_emb not in (1, 2):
            raise ValueError("n_pre_emb must be 1 (TED) or 2 (BEAT)")
        if audio_len is None:
            audio_len = 36267 if n_pre_emb == 1 else 36266
        self.dims = RagDims("ted" if n_pre_emb == 1 else "beat", njoints, nfeats, n_pre_emb, audio_len,
                            latent_dim, num_layers)
        self._build_parameters()
        self._engine = None
        self._engine_sig = None
        self.impl = "auto"

    # ---- parameters --------------------------------------------------------------------
    def _build_parameters(self):
        d, S, dm = self.latent_dim, self.dims.seq_len, self.dims

        def lin(shape, fan_in):          # torch's default Linear/Conv init range
            b = 1.0 / math.sqrt(fan_in)
            return torch.empty(shape).uniform_(-b, b)

        for l in range(self.num_layers):
            p = "backbone.mlps.%d." % l
            _attach(self, p + "block1.0.alpha", torch.ones(1, 1, d))
            _attach(self, p + "block1.0.beta", torch.zeros(1, 1, d))
            _attach(self, p + "block1.1.weight", lin((S, S, 1), S))
            _attach(self, p + "block1.1.bias", lin((S,), S))
            _attach(self, p + "block2.0.alpha", torch.ones(1, 1, d))
            _attach(self, p + "block2.0.beta", torch.zeros(1, 1, d))
            w = torch.empty(d, d)
            nn.init.xavier_uniform_(w, gain=1e-8)       # mlp_module.py:63-65
            _attach(self, p + "block2.1.weight", w)
            _attach(self, p + "block2.1.bias", torch.zeros(d))
        _attach(self, "backbone.sequence_pos_encoder.pe", positional_table(d), buffer=True)
        # the timestep embedder shares the backbone's positional encoder (mlp_module.py:82-83)
        _attach(self, "backbone.embed_timestep.sequence_pos_encoder.pe", None,
                share=self.backbone.sequence_pos_encoder)
        for i in (0, 2):
            _attach(self, "backbone.embed_timestep.time_embed.%d.weight" % i, lin((d, d), d))
            _attach(self, "backbone.embed_timestep.time_embed.%d.bias" % i, lin((d,), d))
        _attach(self, "input_mapping.weight", lin((d, dm.in_feats), dm.in_feats))
        _attach(self, "input_mapping.bias", lin((d,), dm.in_feats))
        _attach(self, "sequence_pos_encoder.pe", positional_table(d), buffer=True)
        _attach(self, "speaker_embedding.weight", torch.full((N_SPEAKERS, SPEAKER_DIM), 1e-6))   # RAG.py:67
        for n in ("speaker_mu", "speaker_logvar"):
            _attach(self, n + ".weight", lin((d, SPEAKER_DIM), SPEAKER_DIM))
            _attach(self, n + ".bias", lin((d,), SPEAKER_DIM))
        if dm.n_pre_emb == 2:
            _attach(self, "emotion_embedding.weight", torch.full((N_EMOTIONS, d), 1e-6))
        for idx, (co, ci) in zip((0, 3, 6, 9), ((32, 1), (64, 32), (128, 64), (256, 128))):
            p = "audio_encoder.feat_extractor.%d." % idx
            _attach(self, p + "weight", lin((co, ci, 15), ci * 15))
            _attach(self, p + "bias", lin((co,), ci * 15))
        _attach(self, "output_process.poseFinal.weight", lin((dm.jd, d), d))
        _attach(self, "output_process.poseFinal.bias", lin((dm.jd,), d))

    def parameters_wo_clip(self):
        return [p for name, p in self.named_parameters() if not name.startswith('clip_model.')]

    # ---- engine ------------------------------------------------------------------------
    def _weights_signature(self):
        sig = [str(self.input_mapping.weight.device)]
        for t in list(self.parameters()) + list(self.buffers()):
            sig.append((t.data_ptr(), t._version))
        return tuple(sig)

    def engine(self, batch=None):
        """The C-ABI engine for the device the parameters live on; (re)uploads the
        weights when they changed (load_state_dict, .to(), in-place edits)."""
        dev = self.input_mapping.weight.device
        need = max(int(batch or 1), 1)
        if self._engine is None or self._engine.device != dev or need > self._engine.max_batch:
            cap = max(need, self._engine.max_batch if self._engine is not None else 0)
            self._engine = None
            self._engine = _cabi.Engine(self.dims, dev, max_batch=cap, n_speakers=N_SPEAKERS,
                                        n_emotions=N_EMOTIONS if self.dims.n_pre_emb == 2 else 0)
            self._engine_sig = None
        sig = self._weights_signature()
        if sig != self._engine_sig:
            self._engine.load_state_dict(self.state_dict())
            self._engine_sig = sig
        self._engine.set_impl(self.impl)
        return self._engine

    def set_impl(self, impl):
        """'auto' | 'simt' | 'tc_bf16x3' | 'tc_bf16' (include/livelyspeaker_b200.h LS_IMPL_*)."""
        if impl not in _cabi.IMPL_NAMES:
            raise ValueError(impl)
        self.impl = impl

    # ---- forward -----------------------------------------------------------------------
    def mask_cond(self, cond, force_mask=False):
        """RAG.py:80-96 (the kernels apply the mask themselves; this mirror is for callers that use it directly)."""
        bs = cond.shape[0]
        if force_mask:
            return torch.zeros_like(cond)
        if self.training and self.cond_mask_prob > 0.:
            mask = torch.bernoulli(torch.ones(bs, device=cond.device) * self.cond_mask_prob).view(bs, 1)
            return (cond.flatten(1) * (1. - mask)).reshape(cond.shape)
        return cond

    def forward(self, x, timesteps, y=None):
        """x: [B, njoints, nfeats, 34] (x_t); timesteps: [B] int; y: the cond dict.
        Returns {'output' [B,J,D,34], 'z_mu', 'z_logvar' [B,1,512]} (RAG.py:98-133)."""
        bs, njoints, nfeats, nframes = x.shape
        if (njoints, nfeats, nframes) != (self.njoints, self.nfeats, N_FRAMES):
            raise ValueError("x must be [B,%d,%d,%d]" % (self.njoints, self.nfeats, N_FRAMES))
        eng = self.engine(bs)
        eng.set_cond(y)
        # reparameterize (RAG.py:10-13) draws even in eval mode; keep torch's generator order
        force_mask = bool(y.get('uncond', False))
        if self.training and self.cond_mask_prob > 0. and not force_mask:
            # training mode (RAG.py:84-93): one Bernoulli draw per clip BEFORE the style draw, 1 = null condition.
            # Forward values only - there is no backward with respect to the weights.
            drop = torch.bernoulli(torch.ones(bs, device=eng.device) * self.cond_mask_prob)
            eps = torch.randn(bs, 1, self.latent_dim, device=eng.device)
            out, z_mu, z_logvar = eng.model_forward_train(x, timesteps, drop, eps)
            return {'output': out, 'z_mu': z_mu, 'z_logvar': z_logvar}
        eps = torch.randn(bs, 1, self.latent_dim, device=eng.device)
        out, z_mu, z_logvar = eng.model_forward(x, timesteps, force_mask, eps)
        return {'output': out, 'z_mu': z_mu, 'z_logvar': z_logvar}
