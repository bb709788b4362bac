"""Classifier-free-guidance sampling wrapper (scripts/model/cfg_sampler.py:8-31).

Same attribute surface and ``forward(x, timesteps, y)`` contract as the reference's
``ClassifierFreeSampleModel``; both denoiser passes and the guidance combine run
inside one C-ABI call (``ls_cfg_forward``) instead of two module calls plus a
``deepcopy`` of the whole cond dict per step.
"""
import torch
import torch.nn as nn


class ClassifierFreeSampleModel(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model
        self.translation = model.translation
        self.njoints = model.njoints
        self.nfeats = model.nfeats
        self.data_rep = model.data_rep
        self.cond_mode = model.cond_mode
        # Optional source of the two style draws of a call (an object with randn(shape, device), e.g. the
        # ReplayNoise a test shares with the diffusion); None = torch's generator, like the reference.
        self.noise_source = None

    def forward(self, x, timesteps, y=None):
        # cfg_sampler.py:25: a model trained without condition dropout falls through and
        # returns None in the reference; kept, because callers can observe it.
        if not self.model.cond_mask_prob > 0:
            return None
        inner = self.model
        if inner.training:
            raise NotImplementedError("sampling wrapper: call .eval() first")
        bs = x.shape[0]
        eng = inner.engine(bs)
        eng.set_cond(y)
        # draw order of the reference: cond pass first, then the uncond pass
        if self.noise_source is not None:
            eps_c = self.noise_source.randn((bs, 1, inner.latent_dim), eng.device)
            eps_u = self.noise_source.randn((bs, 1, inner.latent_dim), eng.device)
        else:
            eps_c = torch.randn(bs, 1, inner.latent_dim, device=eng.device)
            eps_u = torch.randn(bs, 1, inner.latent_dim, device=eng.device)
        return eng.cfg_forward(x, timesteps, eps_c, eps_u, y['scale'])
