"""Cases of the *_with_grad fixtures (grad_ted.npz / grad_beat.npz): shared by the generator (make_golden_grad.py, runs
the reference) and the tests (never import the reference).

cond_fn has the *_with_grad signature cond_fn(x, t, p_mean_var, **model_kwargs) (gaussian_diffusion.py:453, 497) and
differentiates a function of p_mean_var['pred_xstart'] with respect to x - the one thing the plain cond_fn route cannot
do, and the reason these samplers need a backward pass of the denoiser."""
import torch

from livelyspeaker_b200 import synthetic

# tag: (respacing, ddim, seed, first spaced index, n chained steps, kwargs)
CASES = {
    "anc_hi": ("ddim100", False, 501, 99, 3, {}),
    "anc_lo_clip": ("ddim100", False, 502, 2, 3, {"clip_denoised": True}),       # ends with the t == 0 step
    "ddim_mid": ("ddim100", True, 503, 60, 3, {}),
    "ddim_eta_lo": ("ddim100", True, 504, 2, 3, {"eta": 0.5}),
    "anc_nocond": ("ddim100", False, 505, 40, 2, {"no_cond_fn": True}),           # cond_fn=None: plain step under enable_grad
}
VJP_T = (999, 500, 3)       # ORIGINAL timesteps of the bare vector-Jacobian checks (one per clip, cycled)


def make_cond_fn(dims, B, weight=0.05):
    g = torch.Generator().manual_seed(78)
    shape = (B, dims.njoints, dims.nfeats, synthetic.N_FRAMES)
    target = 0.2 * torch.randn(shape, generator=g)

    def cond_fn(x, t, p_mean_var, y=None):
        # grad_x of -w(t) * |pred_xstart(x) - target|^2  +  a direct term in x
        w = weight * (1.0 + t.float().view(-1, 1, 1, 1) / 100.0)
        with torch.enable_grad():
            loss = -(w * (p_mean_var["pred_xstart"] - target.to(x.device)) ** 2).sum() - 0.01 * (x ** 2).sum()
            return torch.autograd.grad(loss, x)[0]
    return cond_fn, target
