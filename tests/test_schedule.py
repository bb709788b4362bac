"""Index / schedule path: bit-exact against the reference's own tables (fixtures written
by tests/golden/make_golden.py from /root/reference) - SURVEY.md section 8 rows a1-a3."""
import json
import types

import numpy as np
import pytest
import torch

import livelyspeaker_b200 as ls
from livelyspeaker_b200.respace import _WrappedModel
from oracle import schedule_oracle

TABLES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
          "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
          "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
          "posterior_mean_coef1", "posterior_mean_coef2"]


def _args(steps, name):
    return types.SimpleNamespace(diffusion_steps=steps, noise_schedule=name, sigma_small=True, lambda_vel=1.0,
                                 lambda_rcxyz=0.0, lambda_fc=0.0)


def test_tables_bit_exact_product_and_oracle(golden_schedule):
    cases = json.loads(str(golden_schedule["cases"]))
    assert len(cases) >= 9
    for k, (name, steps, spec) in enumerate(cases):
        d = ls.create_gaussian_diffusion(_args(steps, name), spec)
        tab, tmap = schedule_oracle.build(name, steps, spec)
        want_map = golden_schedule["c%d_timestep_map" % k]
        assert np.array_equal(np.array(d.timestep_map, dtype=np.int64), want_map)
        assert np.array_equal(np.array(tmap, dtype=np.int64), want_map)
        assert d.num_timesteps == int(golden_schedule["c%d_num_timesteps" % k])
        for key in TABLES:
            want = golden_schedule["c%d_%s" % (k, key)]
            got = getattr(d, key)
            assert got.dtype == np.float64 and np.array_equal(got, want), (name, steps, spec, key)
            assert np.array_equal(tab[key], want), ("oracle", name, steps, spec, key)


def test_known_answers():
    # SURVEY.md appendix B
    assert ls.space_timesteps(1000, "ddim100") == set(range(0, 1000, 10))
    assert ls.space_timesteps(1000, [1000]) == set(range(1000))
    assert ls.space_timesteps(300, "10,15,20") == schedule_oracle.kept_timesteps(300, "10,15,20")
    assert ls.space_timesteps(300, [10, 15, 20]) == ls.space_timesteps(300, "10,15,20")
    d = ls.create_gaussian_diffusion(_args(1000, "cosine"), "")
    assert d.betas[0] == 4.128422482196914e-05 and d.betas[-1] == 0.999
    assert d.alphas_cumprod[-1] == 2.4287669070348567e-09
    d = ls.create_gaussian_diffusion(_args(1000, "cosine"), "ddim100")
    assert d.timestep_map == list(range(0, 1000, 10)) and d.num_timesteps == 100
    assert d.posterior_mean_coef1[99] == 0.022966299813352853
    assert d.posterior_mean_coef2[99] == 0.47341577964733306
    assert d.posterior_variance[99] == 0.775045058990057
    assert d.model_mean_type == ls.ModelMeanType.START_X and d.model_var_type == ls.ModelVarType.FIXED_SMALL
    assert d.loss_type == ls.LossType.HUBER and d.rescale_timesteps is False


def test_space_timesteps_errors():
    with pytest.raises(ValueError):
        ls.space_timesteps(1000, "ddim333")
    with pytest.raises(ValueError):
        ls.space_timesteps(10, [20])
    assert ls.space_timesteps(10, [1]) == {0}
    assert ls.space_timesteps(7, "3,2") == schedule_oracle.kept_timesteps(7, "3,2")


def test_wrapped_model_index_map_bit_exact():
    d = ls.create_gaussian_diffusion(_args(1000, "cosine"), "ddim50")
    seen = {}

    def model(x, ts, **kw):
        seen["ts"] = ts
        return x

    ts = torch.tensor([49, 0, 7, 7, 23])
    _WrappedModel(model, d.timestep_map, False, 1000)(torch.zeros(5), ts)
    assert seen["ts"].dtype == torch.int64
    assert seen["ts"].tolist() == [d.timestep_map[i] for i in ts.tolist()]
    assert [d._model_timestep(i) for i in range(50)] == d.timestep_map


def test_step_params_match_torch_fp32_arithmetic():
    """ls_step_params scalars must be the fp32 numbers torch computes on the gathered tensors."""
    d = ls.create_gaussian_diffusion(_args(1000, "cosine"), "ddim100")
    for i in (99, 57, 1, 0):
        for eta in (0.0, 0.5, 1.0):
            p = d.step_params(i, ddim=True, eta=eta, clip_denoised=False)
            ab = torch.from_numpy(d.alphas_cumprod)[i].float()
            abp = torch.from_numpy(d.alphas_cumprod_prev)[i].float()
            sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
            want = [torch.from_numpy(d.sqrt_recip_alphas_cumprod)[i].float(),
                    torch.from_numpy(d.sqrt_recipm1_alphas_cumprod)[i].float(), torch.sqrt(abp),
                    torch.sqrt(1 - abp - sigma ** 2), sigma]
            got = [p.c[k] for k in range(5)]
            assert got == [float(w) for w in want], (i, eta)
            assert (p.mode, p.t_model, p.add_noise) == (1, d.timestep_map[i], int(i != 0))
        p = d.step_params(i, ddim=False, clip_denoised=True)
        assert p.c[0] == float(np.float32(d.posterior_mean_coef1[i]))
        assert p.c[1] == float(np.float32(d.posterior_mean_coef2[i]))
        assert p.c[2] == float(np.float32(d.posterior_log_variance_clipped[i]))
        assert (p.mode, p.clip_denoised) == (0, 1)
