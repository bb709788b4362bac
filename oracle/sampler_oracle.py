"""Oracle (test infrastructure): ancestral / DDIM sampling loops.

torch-CPU fp32 restatement of the reference's sampling loop with the random
draws made explicit.  ``NoiseTape`` performs the draws in the reference's
order and records them so the CUDA path can replay the
identical numbers.

Reference:
  scripts/diffusion/gaussian_diffusion.py:240-258    q_sample
  scripts/diffusion/gaussian_diffusion.py:260-282    q_posterior_mean_variance
  scripts/diffusion/gaussian_diffusion.py:284-399    p_mean_variance (START_X, FIXED_SMALL; inpainting blend :314-320)
  scripts/diffusion/gaussian_diffusion.py:429-481    condition_mean / condition_score (cond_fn hooks)
  scripts/diffusion/gaussian_diffusion.py:507-558    p_sample
  scripts/diffusion/gaussian_diffusion.py:673-743    p_sample_loop_progressive
  scripts/diffusion/gaussian_diffusion.py:745-798    ddim_sample
  scripts/diffusion/gaussian_diffusion.py:945-1014   ddim_sample_loop_progressive
  scripts/diffusion/gaussian_diffusion.py:1651-1664  _extract_into_tensor (fp64 -> fp32 cast)
  scripts/diffusion/respace.py:118-130               _WrappedModel (spaced t -> original t)
"""
import numpy as np
import torch

from . import rag_oracle


class NoiseTape:
    """Draws N(0,1) tensors from a torch generator in call order and keeps them."""

    def __init__(self, seed=None, device="cpu", replay=None):
        self.replay = list(replay) if replay is not None else None
        self.record = []
        self.gen = None
        self.device = device
        if self.replay is None:
            self.gen = torch.Generator(device=device)
            self.gen.manual_seed(seed)

    def draw(self, *shape):
        """th.randn(*shape): a fresh contiguous tensor."""
        if self.replay is not None:
            t = self.replay[len(self.record)]
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        else:
            t = torch.randn(*shape, generator=self.gen, device=self.device)
        self.record.append(t)
        return t

    def draw_like(self, x):
        """th.randn_like(x): same STRIDES as x.  This matters: from the second step on
        x_t is a dense but permuted tensor (memory order [F,B,J,D], inherited from
        OutputProcess' permute, RAG.py:209-210), and torch fills such a tensor in
        memory order - on CPU through a different code path that also consumes the
        generator differently.  Drawing on an empty_like(x) reproduces both effects."""
        if self.replay is not None:
            t = self.replay[len(self.record)]
            assert tuple(t.shape) == tuple(x.shape), (t.shape, x.shape)
        else:
            t = torch.empty_like(x).normal_(generator=self.gen)
        self.record.append(t)
        return t


def _pick(table, i):
    """_extract_into_tensor for a batch-uniform index: fp64 entry -> fp32 scalar tensor."""
    return torch.from_numpy(np.asarray(table))[i].float()


def q_sample(tab, x0, i, noise):
    return _pick(tab["sqrt_alphas_cumprod"], i) * x0 + _pick(tab["sqrt_one_minus_alphas_cumprod"], i) * noise


def _model_x0(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised, hooks=None):
    """p_mean_variance up to pred_xstart (gaussian_diffusion.py:308-372): model call (two style draws), optional
    inpainting blend (:314-320; TED re-noises the known motion to level i-1 with a fresh randn_like draw unless
    i == 0, the BEAT tree blends it as is, scripts_beat/...:319), optional denoised_fn, optional clamp."""
    B = x.shape[0]
    t_orig = torch.full((B,), tmap[i], dtype=torch.long)
    e_c = tape.draw(B, 1, sd["speaker_mu.weight"].shape[0])
    e_u = tape.draw(B, 1, sd["speaker_mu.weight"].shape[0])
    x0 = rag_oracle.cfg_forward(sd, x, t_orig, y, e_c, e_u, nj, nf)
    hooks = hooks or {}
    if hooks.get("inpaint") is not None:
        mask, motion, noised = hooks["inpaint"]
        if noised and i > 0:
            motion = q_sample(tab, motion, i - 1, tape.draw_like(motion))
        x0 = (x0 * ~mask) + (motion * mask)
    if hooks.get("denoised_fn") is not None:
        x0 = hooks["denoised_fn"](x0)
    if clip_denoised:
        x0 = x0.clamp(-1, 1)
    return x0


def _cond_grad(hooks, tmap, x, i, y):
    """cond_fn as SpacedDiffusion calls it: through _WrappedModel, i.e. with the ORIGINAL timestep (respace.py:100-104)."""
    t_orig = torch.full((x.shape[0],), tmap[i], dtype=torch.long)
    return hooks["cond_fn"](x, t_orig, y=y)


def p_sample_step(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised=False, const_noise=False, hooks=None):
    """One ancestral step.  Returns (sample, pred_xstart)."""
    x0 = _model_x0(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised, hooks)
    mean = _pick(tab["posterior_mean_coef1"], i) * x0 + _pick(tab["posterior_mean_coef2"], i) * x
    log_var = _pick(tab["posterior_log_variance_clipped"], i)
    noise = tape.draw_like(x)      # gaussian_diffusion.py:543 / :787 randn_like(x)
    if const_noise:
        noise = noise[[0]].repeat(x.shape[0], 1, 1, 1)
    if hooks and hooks.get("cond_fn") is not None:       # condition_mean (:429-442), FIXED_SMALL variance
        mean = mean.float() + _pick(tab["posterior_variance"], i) * _cond_grad(hooks, tmap, x, i, y).float()
    nz = 0.0 if i == 0 else 1.0
    return mean + nz * torch.exp(0.5 * log_var) * noise, x0


def ddim_step(sd, tab, tmap, x, i, y, tape, nj, nf, eta=0.0, clip_denoised=False, const_noise=False, hooks=None):
    """One DDIM step.  Returns (sample, pred_xstart)."""
    x0 = _model_x0(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised, hooks)
    x0_ret = x0
    recip, recipm1 = _pick(tab["sqrt_recip_alphas_cumprod"], i), _pick(tab["sqrt_recipm1_alphas_cumprod"], i)
    ab = _pick(tab["alphas_cumprod"], i)
    if hooks and hooks.get("cond_fn") is not None:       # condition_score (:456-481)
        eps = (recip * x - x0) / recipm1
        eps = eps - (1 - ab).sqrt() * _cond_grad(hooks, tmap, x, i, y)
        x0 = recip * x - recipm1 * eps
    eps = (recip * x - x0) / recipm1
    ab_prev = _pick(tab["alphas_cumprod_prev"], i)
    sigma = eta * torch.sqrt((1 - ab_prev) / (1 - ab)) * torch.sqrt(1 - ab / ab_prev)
    noise = tape.draw_like(x)      # gaussian_diffusion.py:543 / :787 randn_like(x)
    if const_noise:
        noise = noise[[0]].repeat(x.shape[0], 1, 1, 1)
    mean = x0 * torch.sqrt(ab_prev) + torch.sqrt(1 - ab_prev - sigma ** 2) * eps
    nz = 0.0 if i == 0 else 1.0
    return mean + nz * sigma * noise, x0_ret


def _p_mean_var_grad(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised, hooks=None):
    """p_mean_variance under enable_grad with x.requires_grad_() (gaussian_diffusion.py:590-599, 819-828): the dict the
    reference hands to cond_fn(x, t, p_mean_var, **model_kwargs), still attached to x's autograd graph."""
    x0 = _model_x0(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised, hooks)
    mean = _pick(tab["posterior_mean_coef1"], i) * x0 + _pick(tab["posterior_mean_coef2"], i) * x
    shape = x.shape
    return {"mean": mean, "variance": _pick(tab["posterior_variance"], i).expand(shape),
            "log_variance": _pick(tab["posterior_log_variance_clipped"], i).expand(shape), "pred_xstart": x0}


def p_sample_with_grad_step(sd, tab, tmap, x, i, y, tape, nj, nf, cond_fn, clip_denoised=False, hooks=None):
    """p_sample_with_grad + condition_mean_with_grad (gaussian_diffusion.py:560-606, 444-456).  cond_fn is called with
    the SPACED index tensor (SpacedDiffusion wraps condition_mean / condition_score only, respace.py:100-104) and may
    differentiate p_mean_var with respect to x.  Returns (sample, pred_xstart)."""
    t = torch.full((x.shape[0],), i, dtype=torch.long)
    with torch.enable_grad():
        x = x.detach().requires_grad_()
        out = _p_mean_var_grad(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised, hooks)
        noise = tape.draw_like(x)
        mean = out["mean"]
        if cond_fn is not None:
            mean = mean.float() + out["variance"] * cond_fn(x, t, out, y=y).float()
    nz = 0.0 if i == 0 else 1.0
    sample = mean + nz * torch.exp(0.5 * out["log_variance"]) * noise
    return sample.detach(), out["pred_xstart"].detach()


def ddim_sample_with_grad_step(sd, tab, tmap, x, i, y, tape, nj, nf, cond_fn, eta=0.0, clip_denoised=False, hooks=None):
    """ddim_sample_with_grad + condition_score_with_grad (gaussian_diffusion.py:800-855, 483-505)."""
    t = torch.full((x.shape[0],), i, dtype=torch.long)
    recip, recipm1 = _pick(tab["sqrt_recip_alphas_cumprod"], i), _pick(tab["sqrt_recipm1_alphas_cumprod"], i)
    ab, ab_prev = _pick(tab["alphas_cumprod"], i), _pick(tab["alphas_cumprod_prev"], i)
    with torch.enable_grad():
        x = x.detach().requires_grad_()
        out = _p_mean_var_grad(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised, hooks)
        x0_orig = out["pred_xstart"]
        x0 = x0_orig
        if cond_fn is not None:
            eps = (recip * x - x0) / recipm1
            eps = eps - (1 - ab).sqrt() * cond_fn(x, t, out, y=y)
            x0 = recip * x - recipm1 * eps
    x0 = x0.detach()
    x = x.detach()
    eps = (recip * x - x0) / recipm1
    sigma = eta * torch.sqrt((1 - ab_prev) / (1 - ab)) * torch.sqrt(1 - ab / ab_prev)
    noise = tape.draw_like(x)
    mean = x0 * torch.sqrt(ab_prev) + torch.sqrt(1 - ab_prev - sigma ** 2) * eps
    nz = 0.0 if i == 0 else 1.0
    return mean + nz * sigma * noise, x0_orig.detach()


def ddim_reverse_step(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised=False):
    """ddim_reverse_sample (gaussian_diffusion.py:857-893): x_i -> x_{i+1} along the deterministic DDIM ODE.  One model
    call (two style draws), no step noise.  Returns (sample, pred_xstart)."""
    x0 = _model_x0(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised)
    eps = (_pick(tab["sqrt_recip_alphas_cumprod"], i) * x - x0) / _pick(tab["sqrt_recipm1_alphas_cumprod"], i)
    ab_next = _pick(tab["alphas_cumprod_next"], i)
    return x0 * torch.sqrt(ab_next) + torch.sqrt(1 - ab_next) * eps, x0


def normal_kl(mean1, logvar1, mean2, logvar2):
    """scripts/diffusion/losses.py:12-43 (tensor arguments)."""
    return 0.5 * (-1.0 + logvar2 - logvar1 + torch.exp(logvar1 - logvar2) + ((mean1 - mean2) ** 2) * torch.exp(-logvar2))


def _approx_cdf(x):
    """losses.py:46-51."""
    return 0.5 * (1.0 + torch.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * torch.pow(x, 3))))


def discretized_gaussian_log_likelihood(x, means, log_scales):
    """losses.py:54-77: log-probability of the 1/255-wide bin around x (data assumed rescaled to [-1, 1])."""
    centered = x - means
    inv_stdv = torch.exp(-log_scales)
    cdf_plus = _approx_cdf(inv_stdv * (centered + 1.0 / 255.0))
    cdf_min = _approx_cdf(inv_stdv * (centered - 1.0 / 255.0))
    log_cdf_plus = torch.log(cdf_plus.clamp(min=1e-12))
    log_one_minus_cdf_min = torch.log((1.0 - cdf_min).clamp(min=1e-12))
    mid = torch.log((cdf_plus - cdf_min).clamp(min=1e-12))
    return torch.where(x < -0.999, log_cdf_plus, torch.where(x > 0.999, log_one_minus_cdf_min, mid))


def _mean_flat(v):
    return v.mean(dim=list(range(1, v.dim())))


def vb_terms_bpd(sd, tab, tmap, x_start, x_t, i, y, tape, nj, nf, clip_denoised=True):
    """_vb_terms_bpd (gaussian_diffusion.py:1213-1247) at the batch-uniform spaced index i, fixed small variance.
    One model call (two style draws).  Returns (output [N] in bits, pred_xstart)."""
    c1, c2 = _pick(tab["posterior_mean_coef1"], i), _pick(tab["posterior_mean_coef2"], i)
    log_var = _pick(tab["posterior_log_variance_clipped"], i) * torch.ones_like(x_start)
    true_mean = c1 * x_start + c2 * x_t
    x0 = _model_x0(sd, tab, tmap, x_t, i, y, tape, nj, nf, clip_denoised)
    mean = c1 * x0 + c2 * x_t
    if i == 0:
        out = _mean_flat(-discretized_gaussian_log_likelihood(x_start, mean, 0.5 * log_var)) / np.log(2.0)
    else:
        out = _mean_flat(normal_kl(true_mean, log_var, mean, log_var)) / np.log(2.0)
    return out, x0


def prior_bpd(tab, x_start):
    """_prior_bpd (:1573-1590): KL(q(x_T | x_0) || N(0, I)) in bits per dimension."""
    T = len(tab["betas"])
    qt_mean = _pick(tab["sqrt_alphas_cumprod"], T - 1) * x_start
    qt_log_var = _pick(tab["log_one_minus_alphas_cumprod"], T - 1) * torch.ones_like(x_start)
    zero = torch.zeros(())
    return _mean_flat(normal_kl(qt_mean, qt_log_var, zero, zero)) / np.log(2.0)


def calc_bpd_loop(sd, tab, tmap, x_start, y, tape, nj, nf, clip_denoised=True):
    """calc_bpd_loop (:1592-1645): per step one randn_like draw, then the model's two style draws."""
    T = len(tab["betas"])
    vb, xstart_mse, mse = [], [], []
    for i in range(T - 1, -1, -1):
        noise = tape.draw_like(x_start)
        x_t = q_sample(tab, x_start, i, noise)
        out, x0 = vb_terms_bpd(sd, tab, tmap, x_start, x_t, i, y, tape, nj, nf, clip_denoised)
        vb.append(out)
        xstart_mse.append(_mean_flat((x0 - x_start) ** 2))
        eps = (_pick(tab["sqrt_recip_alphas_cumprod"], i) * x_t - x0) / _pick(tab["sqrt_recipm1_alphas_cumprod"], i)
        mse.append(_mean_flat((eps - noise) ** 2))
    vb, xstart_mse, mse = torch.stack(vb, dim=1), torch.stack(xstart_mse, dim=1), torch.stack(mse, dim=1)
    prior = prior_bpd(tab, x_start)
    return {"total_bpd": vb.sum(dim=1) + prior, "prior_bpd": prior, "vb": vb, "xstart_mse": xstart_mse, "mse": mse}


def training_losses(sd, tab, tmap, x_start, t_idx, y, noise, style_eps, cond_drop, nj, nf, lambda_vel=1.0):
    """GaussianDiffusion.training_losses, LossType.HUBER (gaussian_diffusion.py:1249-1401): q_sample at the per-clip
    spaced indices t_idx, RAG.forward in training mode (cond_drop = the Bernoulli mask of mask_cond, RAG.py:84-93;
    style_eps = reparameterize's draw), compute_huber (:21-24) of the sample and of its frame differences, kld."""
    import torch.nn.functional as F
    a = torch.from_numpy(np.asarray(tab["sqrt_alphas_cumprod"]))[t_idx].float().view(-1, 1, 1, 1)
    b = torch.from_numpy(np.asarray(tab["sqrt_one_minus_alphas_cumprod"]))[t_idx].float().view(-1, 1, 1, 1)
    x_t = a * x_start + b * noise
    t_orig = torch.as_tensor(np.asarray(tmap))[t_idx].long()
    yy = dict(y)
    yy["cond_drop"] = cond_drop
    out = rag_oracle.rag_forward(sd, x_t, t_orig, yy, style_eps, nj, nf)

    def huber(p, q):
        return F.smooth_l1_loss(p / 0.1, q / 0.1) * 0.1
    o = out["output"]
    terms = {"rot_mse": huber(x_start, o),
             "vel_mse": huber(x_start[..., 1:] - x_start[..., :-1], o[..., 1:] - o[..., :-1]),
             "kld": -0.5 * torch.mean(1 + out["z_logvar"] - out["z_mu"].pow(2) - out["z_logvar"].exp())}
    terms["loss"] = terms["rot_mse"] + lambda_vel * terms["vel_mse"]
    return terms, o


def sample_loop(sd, tab, tmap, shape, y, tape, ddim=False, eta=0.0, clip_denoised=False,
                skip_timesteps=0, init_image=None, const_noise=False, noise=None, trace=None, hooks=None,
                const_noise_init=True):
    """p_sample_loop / ddim_sample_loop.  ``trace`` (a list) receives
    (i, sample, pred_xstart) per step when given.  ``hooks``: {'inpaint': (mask, motion, noised), 'cond_fn': f,
    'denoised_fn': g}.  ``const_noise_init`` False = the BEAT tree, whose p_sample_loop_progressive does not repeat
    the first clip's initial noise (scripts_beat/diffusion/gaussian_diffusion.py:700-704 vs scripts/...:704-708)."""
    B, nj, nf, _ = shape
    n_t = len(tab["betas"])
    img = noise if noise is not None else tape.draw(*shape)
    if noise is None and const_noise and const_noise_init:
        img = img[[0]].repeat(B, 1, 1, 1)
    if skip_timesteps and init_image is None:
        init_image = torch.zeros_like(img)
    indices = list(range(n_t - skip_timesteps))[::-1]
    if init_image is not None:
        img = q_sample(tab, init_image, indices[0], img)
    step = ddim_step if ddim else p_sample_step
    for i in indices:
        kw = {"eta": eta} if ddim else {}
        img, x0 = step(sd, tab, tmap, img, i, y, tape, nj, nf, clip_denoised=clip_denoised,
                       const_noise=const_noise, hooks=hooks, **kw)
        if trace is not None:
            trace.append((i, img, x0))
    return img


def plms_loop(sd, tab, tmap, shape, y, tape, order=2, clip_denoised=False, skip_timesteps=0, init_image=None,
              noise=None, trace=None):
    """plms_sample_loop (gaussian_diffusion.py:1016-1211): pseudo improved Euler on the first step, then
    Adams-Bashforth of the given order on the eps re-derived from the x0 prediction.  The first step calls the
    model twice (second call at index i-1); no step noise is drawn.  order must be 2..4: with order 1 the
    reference dereferences old_out=None on the first step (:1076)."""
    B, nj, nf, _ = shape
    n_t = len(tab["betas"])
    img = noise if noise is not None else tape.draw(*shape)
    if skip_timesteps and init_image is None:
        init_image = torch.zeros_like(img)
    indices = list(range(n_t - skip_timesteps))[::-1]
    if init_image is not None:
        img = q_sample(tab, init_image, indices[0], img)

    def model_eps(x, i):
        x0 = _model_x0(sd, tab, tmap, x, i, y, tape, nj, nf, clip_denoised)
        eps = (_pick(tab["sqrt_recip_alphas_cumprod"], i) * x - x0) / _pick(tab["sqrt_recipm1_alphas_cumprod"], i)
        return eps, x0

    def x0_from_eps(x, i, eps):
        return _pick(tab["sqrt_recip_alphas_cumprod"], i) * x - _pick(tab["sqrt_recipm1_alphas_cumprod"], i) * eps

    old = None
    for i in indices:
        ab_prev = _pick(tab["alphas_cumprod_prev"], i)
        eps, x0 = model_eps(img, i)
        if order > 1 and old is None:
            old = [eps]
            mean = x0 * torch.sqrt(ab_prev) + torch.sqrt(1 - ab_prev) * eps
            eps2, _ = model_eps(mean, i - 1)
            eps_p = (eps + eps2) / 2
        else:
            old.append(eps)
            k = min(order, len(old))
            if k == 1:
                eps_p = old[-1]
            elif k == 2:
                eps_p = (3 * old[-1] - old[-2]) / 2
            elif k == 3:
                eps_p = (23 * old[-1] - 16 * old[-2] + 5 * old[-3]) / 12
            else:
                eps_p = (55 * old[-1] - 59 * old[-2] + 37 * old[-3] - 9 * old[-4]) / 24
        pred = x0_from_eps(img, i, eps_p)
        mean = pred * torch.sqrt(ab_prev) + torch.sqrt(1 - ab_prev) * eps_p
        if len(old) >= order:
            old.pop(0)
        nz = 0.0 if i == 0 else 1.0
        img = mean * nz + x0 * (1 - nz)
        if trace is not None:
            trace.append((i, img, x0))
    return img
