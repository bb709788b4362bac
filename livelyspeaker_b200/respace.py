"""Timestep respacing (scripts/diffusion/respace.py:9-130): `space_timesteps`,
`SpacedDiffusion` (re-derived betas + `timestep_map`) and `_WrappedModel`.
The index path must be bit-exact with the reference (tests/test_schedule.py).
"""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Set of original timesteps to keep.  "ddimN": the integer stride giving exactly N
    steps; otherwise a list / comma string of per-section counts, each section strided
    fractionally with Python's round() (respace.py:9-62)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            wanted = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                picked = range(0, num_timesteps, stride)
                if len(picked) == wanted:
                    return set(picked)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    n_sections = len(section_counts)
    size_per, extra = num_timesteps // n_sections, num_timesteps % n_sections
    steps, start = [], 0
    for idx, count in enumerate(section_counts):
        size = size_per + (1 if idx < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac_stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            steps.append(start + round(cur))
            cur += frac_stride
        start += size
    return set(steps)


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process over a subset of a base process' timesteps."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        self.timestep_map, new_betas, last = [], [], 1.0
        for i, alpha_cumprod in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - alpha_cumprod / last)
                last = alpha_cumprod
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def _model_timestep(self, i):
        return int(self.timestep_map[int(i)])

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def training_losses(self, model, *args, **kwargs):
        return super().training_losses(self._wrap_model(model), *args, **kwargs)

    def condition_mean(self, cond_fn, *args, **kwargs):
        return super().condition_mean(self._wrap_model(cond_fn), *args, **kwargs)

    def condition_score(self, cond_fn, *args, **kwargs):
        return super().condition_score(self._wrap_model(cond_fn), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t   # done by the wrapped model


class _WrappedModel:
    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._maps = {}

    def __call__(self, x, ts, **kwargs):
        key = (ts.device, ts.dtype)
        if key not in self._maps:
            self._maps[key] = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        new_ts = self._maps[key][ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)
