"""Host-side mirror of the reference interface: key set, loader contract, loud failure
without CUDA, and the rule that the product never touches oracle/ or the reference."""
import os
import re
import types

import pytest
import numpy as np
import torch

import livelyspeaker_b200 as ls
from conftest import ROOT
from livelyspeaker_b200 import beat_model_util, synthetic
from livelyspeaker_b200._cabi import LsError


def _args(**kw):
    a = dict(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1, arch='trans_enc',
             emb_trans_dec=False, dataset='humanml', lang_model=None, mlpact='silu', diffusion_steps=1000,
             noise_schedule='cosine', sigma_small=True, lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
    a.update(kw)
    return types.SimpleNamespace(**a)


def test_state_dict_layout_ted():
    model, diffusion = ls.create_model_and_diffusion(_args(), 'ddim100')
    sd = model.state_dict()
    want = synthetic.synth_state_dict(synthetic.TED)
    assert len(sd) == 88 and set(sd) == set(want)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(want[k].shape), k
        assert sd[k].dtype == torch.float32
    assert sum(p.numel() for p in model.parameters()) == 4094267
    assert sum(v.numel() for v in sd.values()) == 11774267
    # the embedder's positional table IS the backbone's (mlp_module.py:82-83)
    assert sd["backbone.sequence_pos_encoder.pe"].data_ptr() == \
        sd["backbone.embed_timestep.sequence_pos_encoder.pe"].data_ptr()
    assert torch.equal(sd["sequence_pos_encoder.pe"], synthetic.positional_table(512))
    # reference init quirks (mlp_module.py:63-65, RAG.py:67)
    assert float(sd["backbone.mlps.0.block2.1.weight"].abs().max()) < 1e-7
    assert float(sd["speaker_embedding.weight"].mean()) == pytest.approx(1e-6)
    assert (model.njoints, model.nfeats, model.cond_mask_prob, model.n_pre_seq) == (9, 3, 0.1, 4)


def test_state_dict_layout_beat():
    model, diffusion = beat_model_util.create_model_and_diffusion(_args(njoints=47), 'ddim100')
    sd = model.state_dict()
    want = synthetic.synth_state_dict(synthetic.BEAT)
    assert set(sd) == set(want) and len(sd) == 89
    for k in sd:
        assert tuple(sd[k].shape) == tuple(want[k].shape), k
    assert sd["backbone.mlps.0.block1.1.weight"].shape == (36, 36, 1)
    assert sd["input_mapping.weight"].shape == (512, 821)
    assert diffusion.dump_key == "sample" and not diffusion.allow_ddim_const_noise


def test_load_model_wo_clip_contract(capsys):
    model, _ = ls.create_model_and_diffusion(_args())
    sd = synthetic.synth_state_dict(synthetic.TED)
    ls.load_model_wo_clip(model, sd)
    assert torch.equal(model.state_dict()["input_mapping.weight"], sd["input_mapping.weight"])
    bad = dict(sd)
    bad["not_a_key"] = torch.zeros(1)
    with pytest.raises(AssertionError):
        ls.load_model_wo_clip(model, bad)
    short = {k: v for k, v in sd.items() if k != "input_mapping.bias"}
    with pytest.raises(AssertionError):
        ls.load_model_wo_clip(model, short)


def test_no_cpu_fallback():
    model, diffusion = ls.create_model_and_diffusion(_args(), 'ddim100')
    cfg = ls.ClassifierFreeSampleModel(model).eval()
    y = synthetic.synth_cond(synthetic.TED, 2)
    with pytest.raises(LsError):
        diffusion.ddim_sample_loop(cfg, (2, 9, 3, 34), model_kwargs={"y": y})
    with pytest.raises(LsError):
        model.eval()(torch.zeros(2, 9, 3, 34), torch.zeros(2, dtype=torch.long), y)


def test_cfg_wrapper_surface_and_none_quirk():
    model, _ = ls.create_model_and_diffusion(_args(cond_mask_prob=0.0))
    cfg = ls.ClassifierFreeSampleModel(model)
    for a in ("model", "njoints", "nfeats", "data_rep", "cond_mode", "translation"):
        assert hasattr(cfg, a)
    # cfg_sampler.py:25 - falls through and returns None
    assert cfg(torch.zeros(1, 9, 3, 34), torch.zeros(1, dtype=torch.long), y={}) is None


def test_training_surface_without_a_gpu():
    """Training mode is a forward-only path since round 2 (ls_model_forward_train / training_losses); like every
    product path it needs the CUDA engine and fails loudly without one - there is no CPU fallback.  The sampling
    wrapper still refuses a model in training mode, and non-HUBER losses are not built."""
    from livelyspeaker_b200 import _cabi
    model, diffusion = ls.create_model_and_diffusion(_args())
    model.train()
    with pytest.raises(_cabi.LsError):
        model(torch.zeros(1, 9, 3, 34), torch.zeros(1, dtype=torch.long), {})
    with pytest.raises(NotImplementedError):
        ls.ClassifierFreeSampleModel(model)(torch.zeros(1, 9, 3, 34), torch.zeros(1, dtype=torch.long), {})
    assert hasattr(diffusion, "training_losses") and hasattr(diffusion, "p_sample_with_grad")
    diffusion.loss_type = ls.gaussian_diffusion.LossType.MSE
    with pytest.raises(NotImplementedError):
        diffusion.training_losses(model, torch.zeros(1, 9, 3, 34), torch.zeros(1, dtype=torch.long),
                                  model_kwargs={"y": {"mask": None}})


def test_generic_route_matches_reference_math_with_a_plain_model():
    """A foreign model (no CUDA engine) goes through the generic torch route: check the
    sampler arithmetic against the oracle's update formulas on CPU."""
    from oracle import sampler_oracle, schedule_oracle

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor(0.7))

        def forward(self, x, t, y=None):
            return self.w * x + 0.01 * t.view(-1, 1, 1, 1).float()

    toy = Toy()
    d = ls.create_gaussian_diffusion(_args(), 'ddim100')
    tab, tmap = schedule_oracle.build("cosine", 1000, "ddim100")
    torch.manual_seed(3)
    got = d.ddim_sample_loop(toy, (2, 9, 3, 34), clip_denoised=False, model_kwargs={"y": {}}, eta=0.3,
                             skip_timesteps=90)
    torch.manual_seed(3)
    x = torch.randn(2, 9, 3, 34)
    x = sampler_oracle.q_sample(tab, torch.zeros_like(x), 9, x)
    for i in range(9, -1, -1):
        with torch.no_grad():
            x0 = toy(x, torch.full((2,), tmap[i]))
        eps = (sampler_oracle._pick(tab["sqrt_recip_alphas_cumprod"], i) * x - x0) / \
            sampler_oracle._pick(tab["sqrt_recipm1_alphas_cumprod"], i)
        ab, abp = sampler_oracle._pick(tab["alphas_cumprod"], i), sampler_oracle._pick(tab["alphas_cumprod_prev"], i)
        sigma = 0.3 * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        nz = torch.randn_like(x)
        x = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps + (0.0 if i == 0 else 1.0) * sigma * nz
    assert torch.allclose(got, x, rtol=1e-6, atol=1e-6)


def test_plms_host_logic_on_cpu_against_the_reference_fixture(golden_plms):
    """PLMS is host logic over p_mean_variance.  Drive the product's plms_sample_loop on CPU with
    a foreign model that evaluates the (reference-pinned) oracle denoiser: the result must be the reference's own
    fixture, draws included (initial noise, then cond / uncond style per model call; no step noise)."""
    from livelyspeaker_b200 import synthetic
    from oracle import rag_oracle
    dims = synthetic.TED
    sd = synthetic.synth_state_dict(dims, seed=1)

    class OracleCfg(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.anchor = torch.nn.Parameter(torch.zeros(1))      # gives the loop its device

        def forward(self, x, t, y=None):
            B = x.shape[0]
            e_c, e_u = torch.randn(B, 1, 512), torch.randn(B, 1, 512)
            return rag_oracle.cfg_forward(sd, x, t, y, e_c, e_u, dims.njoints, dims.nfeats)

    d = ls.create_gaussian_diffusion(_args(), 'ddim100')
    torch.manual_seed(304)
    with torch.no_grad():
        got = d.plms_sample_loop(OracleCfg(), (2, 9, 3, 34), clip_denoised=False,
                                 model_kwargs={"y": synthetic.synth_cond(dims, 2)}, skip_timesteps=80,
                                 init_image=torch.from_numpy(golden_plms["init_image"]), order=2)
    np.testing.assert_allclose(got.numpy(), golden_plms["plms_ddim100_o2_sdedit"], rtol=1e-5, atol=2e-5)
    with pytest.raises(TypeError):       # order 1 dereferences old_out=None on the first step, as in the reference
        d.plms_sample(OracleCfg(), got, torch.zeros(2, dtype=torch.long), order=1, model_kwargs={"y": synthetic.synth_cond(dims, 2)})


def test_sag_decoder_surface_and_no_cpu_path():
    """Same state_dict key set / shapes as the reference's Decoder_TRANSFORMER (the synthetic dict loads strictly
    into the reference module, tests/golden/make_golden_sag.py) and no CPU implementation behind it."""
    from livelyspeaker_b200 import synthetic
    from livelyspeaker_b200._cabi import LsError
    dec = ls.Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False).eval()
    sd = synthetic.synth_sag_state_dict(seed=3)
    assert {k: tuple(v.shape) for k, v in dec.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    dec.load_state_dict(sd, strict=True)
    batch = {"x": torch.zeros(2, 9, 3, 34), "z": torch.zeros(2, 512), "mask": torch.ones(2, 34, dtype=torch.bool)}
    with pytest.raises(LsError):
        dec(batch)
    with pytest.raises(NotImplementedError):
        dec.train()(batch)


def test_product_never_imports_oracle_or_reference():
    pkg = os.path.join(ROOT, "livelyspeaker_b200")
    pat = re.compile(r"^\s*(from|import)\s+(oracle|tests)\b|/root/reference", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), "%s references oracle/ or the reference tree" % f
    bench = open(os.path.join(ROOT, "bench.py")).read() if os.path.exists(os.path.join(ROOT, "bench.py")) else ""
    assert "/root/reference" not in bench


def test_metrics_have_no_cpu_path():
    """The rhythm metric mirrors scripts/test_RAG_ted.py:84-123 on the device only: host tensors raise."""
    from livelyspeaker_b200 import metrics
    from livelyspeaker_b200._cabi import LsError
    with pytest.raises(LsError):
        metrics.motion_beats(torch.zeros(2, 9, 3, 34))
    with pytest.raises(LsError):
        metrics.beat_align_score(torch.zeros(2, 34, dtype=torch.uint8), [[0.1], []])
