"""Oracle (test infrastructure): diffusion schedule + respacing index path.

Pure Python / numpy fp64 restatement.  Everything here must be BIT-exact with
the reference, so the arithmetic is done with the same primitive operations in
the same order (math.cos, np.cumprod, np.sqrt, ...).

Reference:
  scripts/diffusion/gaussian_diffusion.py:26-70   named beta schedules
  scripts/diffusion/gaussian_diffusion.py:167-204 derived fp64 tables
  scripts/diffusion/respace.py:9-62               space_timesteps
  scripts/diffusion/respace.py:74-88              re-derived betas + timestep_map
"""
import math

import numpy as np


def named_betas(name, n_steps, scale=1.0):
    """gaussian_diffusion.py:26-70."""
    if name == "linear":
        s = scale * 1000 / n_steps
        return np.linspace(s * 0.0001, s * 0.02, n_steps, dtype=np.float64)
    if name == "cosine":
        def abar(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        out = []
        for i in range(n_steps):
            a = i / n_steps
            b = (i + 1) / n_steps
            out.append(min(1 - abar(b) / abar(a), 0.999))
        return np.array(out)
    raise NotImplementedError(name)


def kept_timesteps(n_steps, spec):
    """respace.py:9-62.  Returns a python set of original timesteps."""
    if isinstance(spec, str):
        if spec.startswith("ddim"):
            want = int(spec[4:])
            for stride in range(1, n_steps):
                if len(range(0, n_steps, stride)) == want:
                    return set(range(0, n_steps, stride))
            raise ValueError("no integer stride gives %d steps" % want)
        spec = [int(v) for v in spec.split(",")]
    base, extra = divmod(n_steps, len(spec))
    kept, start = [], 0
    for idx, count in enumerate(spec):
        size = base + (1 if idx < extra else 0)
        if size < count:
            raise ValueError("cannot divide section of %d steps into %d" % (size, count))
        step = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += step
        start += size
    return set(kept)


def tables_from_betas(betas):
    """gaussian_diffusion.py:167-204: every derived fp64 table, as a dict."""
    betas = np.array(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    ac_next = np.append(ac[1:], 0.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "alphas_cumprod_next": ac_next,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }


def respaced(base_betas, use_timesteps):
    """respace.py:74-88: (new_betas fp64, timestep_map list)."""
    use = set(use_timesteps)
    ac = tables_from_betas(base_betas)["alphas_cumprod"]
    last = 1.0
    new_betas, tmap = [], []
    for i, a in enumerate(ac):
        if i in use:
            new_betas.append(1 - a / last)
            last = a
            tmap.append(i)
    return np.array(new_betas), tmap


def build(noise_schedule="cosine", diffusion_steps=1000, timestep_respacing=""):
    """model_util.py:40-74 (create_gaussian_diffusion) reduced to its data:
    returns (tables dict for the spaced process, timestep_map)."""
    base = named_betas(noise_schedule, diffusion_steps, 1.0)
    spec = timestep_respacing if timestep_respacing else [diffusion_steps]
    use = sorted(kept_timesteps(diffusion_steps, spec))
    new_betas, tmap = respaced(base, use)
    return tables_from_betas(new_betas), tmap
