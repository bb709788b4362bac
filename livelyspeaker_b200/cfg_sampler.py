"""Classifier-free-guidance sampling wrapper (scripts/model/cfg_sampler.py:8-31).

Same attribute surface and ``forward(x, timesteps, y)`` contract as the reference's
``ClassifierFreeSampleModel``; both denoiser passes and the guidance combine run
inside one C-ABI call (``ls_cfg_forward``) instead of two module calls plus a
``deepcopy`` of the whole cond dict per step.
"""
import torch
import torch.nn as nn


class _CfgGradFn(torch.autograd.Function):
    """The guided denoiser call as an autograd node with respect to x: forward = ``ls_cfg_forward_grad`` (exact-order
    fp32 kernel that keeps every MLPblock's input), backward = ``ls_cfg_backward`` (hand-written vector-Jacobian
    kernel).  This is what the reference gets from torch autograd inside ``p_sample_with_grad`` /
    ``ddim_sample_with_grad`` (gaussian_diffusion.py:560-606, 800-855).  The handle keeps ONE saved forward: backward
    must run before the next differentiable call of the same model (a counter enforces it)."""

    @staticmethod
    def forward(ctx, x, eng, timesteps, eps_c, eps_u, scale):
        out = eng.cfg_forward_grad(x.detach(), timesteps, eps_c, eps_u, scale)
        eng._grad_serial = getattr(eng, "_grad_serial", 0) + 1
        ctx.eng, ctx.scale, ctx.serial = eng, scale, eng._grad_serial
        return out

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.eng._grad_serial != ctx.serial:
            raise RuntimeError("backward through a denoiser call whose saved activations were overwritten by a later "
                               "differentiable call of the same model")
        return ctx.eng.cfg_backward(grad_out, ctx.scale), None, None, None, None, None


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
        if torch.is_grad_enabled() and x.requires_grad:
            # *_with_grad samplers: the call must stay differentiable with respect to x (cond / weights are constants)
            return _CfgGradFn.apply(x, eng, timesteps, eps_c, eps_u, y['scale'])
        return eng.cfg_forward(x, timesteps, eps_c, eps_u, y['scale'])
