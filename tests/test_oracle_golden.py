"""The oracle against the fixtures the REFERENCE produced (tests/golden/make_golden.py).
This is what licenses using oracle/ as the checker in the GPU parity tests."""
import os

import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL
from livelyspeaker_b200 import synthetic
from oracle import rag_oracle, sampler_oracle, schedule_oracle

TIGHT = dict(rtol=1e-5, atol=2e-5)     # oracle vs reference: same fp32 ops, same order


class few_threads:
    """Small-batch oracle loops are thousands of tiny ops: with all cores torch spends its time in thread hand-offs
    (B=1, T=1000: 164 s on 8 threads, 25 s on 2).  Results do not depend on the thread count."""

    def __init__(self, n=2):
        self.n = n

    def __enter__(self):
        self.saved = torch.get_num_threads()
        torch.set_num_threads(self.n)

    def __exit__(self, *exc):
        torch.set_num_threads(self.saved)


def _close(a, b, **kw):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), **(kw or TIGHT))


@pytest.fixture(scope="module")
def ted():
    dims = synthetic.TED
    return dims, synthetic.synth_state_dict(dims, seed=1)


def test_synthetic_weights_are_the_ones_the_fixtures_used(golden_ted, golden_beat, ted):
    for dims, g in ((synthetic.TED, golden_ted), (synthetic.BEAT, golden_beat)):
        sd = synthetic.synth_state_dict(dims, seed=1)
        s = sum(float(v.double().abs().sum()) for v in sd.values())
        assert abs(s - float(g["weights_abs_sum"])) < 1e-6 * s


def test_ops(golden_ted, ted):
    dims, sd = ted
    y = synthetic.synth_cond(dims, 2)
    with torch.no_grad():
        _close(rag_oracle.wav_encoder(sd, y["audio_input"]), golden_ted["wavenc_out"])
        hx = torch.from_numpy(golden_ted["block_in"])
        _close(rag_oracle.ln_spatial(hx, sd["backbone.mlps.3.block1.0.alpha"], sd["backbone.mlps.3.block1.0.beta"]),
               golden_ted["ln_out"])
        tt = torch.from_numpy(golden_ted["temb_t"])
        emb = rag_oracle.timestep_embed(sd, tt)
        _close(emb, golden_ted["temb_out"])
        _close(rag_oracle.mlp_block(sd, 3, hx, emb), golden_ted["block_out"])
        _close(rag_oracle.trans_mlp(sd, hx, tt), golden_ted["backbone_out"])


@pytest.mark.parametrize("name", ["ted", "beat"])
def test_forward_and_cfg(name, golden_ted, golden_beat):
    g = golden_ted if name == "ted" else golden_beat
    dims = synthetic.dims_for(name)
    sd = synthetic.synth_state_dict(dims, seed=1)
    x, t = torch.from_numpy(g["fwd_x"]), torch.from_numpy(g["fwd_t"])
    with torch.no_grad():
        for tag, unc in (("cond", False), ("uncond", True)):
            y = synthetic.synth_cond(dims, 2)
            if unc:
                y["uncond"] = True
            eps = torch.randn(2, 1, 512, generator=torch.Generator().manual_seed(11))
            o = rag_oracle.rag_forward(sd, x, t, y, eps, dims.njoints, dims.nfeats)
            for key in ("output", "z_mu", "z_logvar"):
                _close(o[key], g["fwd_%s_%s" % (tag, key)])
            assert float(y["origin_x"][..., 4:].abs().max()) == 0.0
        tape = sampler_oracle.NoiseTape(seed=12)
        o = rag_oracle.cfg_forward(sd, x, t, synthetic.synth_cond(dims, 2), tape.draw(2, 1, 512),
                                   tape.draw(2, 1, 512), dims.njoints, dims.nfeats)
        _close(o, g["cfg_out"])


@pytest.mark.parametrize("name", ["ted", "beat"])
def test_single_steps(name, golden_ted, golden_beat):
    g = golden_ted if name == "ted" else golden_beat
    dims = synthetic.dims_for(name)
    sd = synthetic.synth_state_dict(dims, seed=1)
    xs = torch.from_numpy(g["step_x"])
    tab_full, map_full = schedule_oracle.build("cosine", 1000, "")
    tab_ddim, map_ddim = schedule_oracle.build("cosine", 1000, "ddim100")
    with torch.no_grad():
        for i in (999, 0):
            tape = sampler_oracle.NoiseTape(seed=100 + i)
            s, p0 = sampler_oracle.p_sample_step(sd, tab_full, map_full, xs, i, synthetic.synth_cond(dims, 2), tape,
                                                 dims.njoints, dims.nfeats)
            _close(s, g["pstep%d_sample" % i])
            _close(p0, g["pstep%d_x0" % i])
        for i, eta in ((57, 0.5), (0, 0.0)):
            tape = sampler_oracle.NoiseTape(seed=200 + i)
            s, p0 = sampler_oracle.ddim_step(sd, tab_ddim, map_ddim, xs, i, synthetic.synth_cond(dims, 2), tape,
                                             dims.njoints, dims.nfeats, eta=eta)
            tag = "dstep%d_eta%d" % (i, int(eta * 10))
            _close(s, g[tag + "_sample"])
            _close(p0, g[tag + "_x0"])


LOOPS = {   # tag: (respacing, steps, ddim, seed, batch, kwargs)
    "ddim100": ("ddim100", 1000, True, 233, 2, {}),
    "anc100": ("100", 1000, False, 234, 2, {}),
    "ddim100_sdedit": ("ddim100", 1000, True, 235, 2, {"skip_timesteps": 80, "init": True}),
    "ddim100_eta05_clip": ("ddim100", 1000, True, 236, 2, {"eta": 0.5, "clip_denoised": True}),
    "anc100_constnoise": ("100", 1000, False, 237, 2, {"const_noise": True}),
    "anc100_skip60": ("100", 1000, False, 238, 2, {"skip_timesteps": 60}),
    "t100_b1": ("", 100, False, 239, 1, {}),
    "anc1000_b1": ("", 1000, False, 240, 1, {}),      # the headline configuration's loop: T=1000 ancestral
}


def run_oracle_loop(tag, g, dims, sd):
    spec, steps, ddim, seed, B, kw = LOOPS[tag]
    kw = dict(kw)
    tab, tmap = schedule_oracle.build("cosine", steps, spec)
    init = torch.from_numpy(g["init_image"]) if kw.pop("init", False) else None
    tape = sampler_oracle.NoiseTape(seed=seed)
    with torch.no_grad(), few_threads():
        out = sampler_oracle.sample_loop(sd, tab, tmap, (B, dims.njoints, dims.nfeats, 34),
                                         synthetic.synth_cond(dims, B), tape, ddim=ddim, init_image=init, **kw)
    return out, tape


@pytest.mark.parametrize("tag", ["ddim100", "anc100", "ddim100_sdedit", "ddim100_eta05_clip", "anc100_constnoise",
                                 "anc100_skip60", "t100_b1", "anc1000_b1"])
def test_whole_loops_ted(tag, golden_ted, ted):
    dims, sd = ted
    out, tape = run_oracle_loop(tag, golden_ted, dims, sd)
    _close(out, golden_ted["loop_" + tag])
    # and within the product tolerance by a wide margin
    np.testing.assert_allclose(out.numpy(), golden_ted["loop_" + tag], rtol=RTOL, atol=ATOL)


PLMS = {   # tag: (respacing, order, seed, kwargs) - tests/golden/make_golden_plms.py
    "ddim100_o2": ("ddim100", 2, 301, {}),
    "ddim100_o3": ("ddim100", 3, 302, {}),
    "ddim50_o4_clip": ("ddim50", 4, 303, {"clip_denoised": True}),
    "ddim100_o2_sdedit": ("ddim100", 2, 304, {"skip_timesteps": 80, "init": True}),
}


def run_oracle_plms(tag, g, dims, sd):
    spec, order, seed, kw = PLMS[tag]
    kw = dict(kw)
    tab, tmap = schedule_oracle.build("cosine", 1000, spec)
    init = torch.from_numpy(g["init_image"]) if kw.pop("init", False) else None
    tape = sampler_oracle.NoiseTape(seed=seed)
    with torch.no_grad(), few_threads():
        out = sampler_oracle.plms_loop(sd, tab, tmap, (2, dims.njoints, dims.nfeats, 34), synthetic.synth_cond(dims, 2),
                                       tape, order=order, init_image=init, **kw)
    return out, tape


@pytest.mark.parametrize("tag", ["ddim100_o3", "ddim50_o4_clip", "ddim100_o2_sdedit"])
def test_plms_loops_ted(tag, golden_plms, ted):
    """PLMS (reference gaussian_diffusion.py:1016-1211): the oracle against the reference's own outputs."""
    dims, sd = ted
    out, _ = run_oracle_plms(tag, golden_plms, dims, sd)
    _close(out, golden_plms["plms_" + tag])


def test_sag_decoder(golden_sag):
    """SAG decoder: the oracle against the reference module's own output
    (tests/golden/make_golden_sag.py: nn.TransformerDecoder of the reference tree)."""
    from oracle import sag_oracle
    sd = synthetic.synth_sag_state_dict(seed=3)
    s = sum(float(v.double().abs().sum()) for v in sd.values())
    assert abs(s - float(golden_sag["weights_abs_sum"])) < 1e-6 * s
    with torch.no_grad():
        out = sag_oracle.decode(sd, torch.from_numpy(golden_sag["x"]), torch.from_numpy(golden_sag["z"]),
                                torch.from_numpy(golden_sag["mask"]))
    _close(out, golden_sag["output"])
    assert float(out[1, :, :, 30:].abs().max()) == 0.0        # padded frames are zeroed


def test_whole_loop_beat(golden_beat):
    dims = synthetic.BEAT
    sd = synthetic.synth_state_dict(dims, seed=1)
    out, _ = run_oracle_loop("ddim100", golden_beat, dims, sd)
    _close(out, golden_beat["loop_ddim100"])


def test_noise_tape_layout_rule():
    """From the 2nd step on the step noise is drawn in [F,B,J,D] memory order (the layout
    OutputProcess' permute gives x_t); the tape must reproduce torch's behaviour."""
    x = torch.empty(34, 2, 9, 3).permute(1, 2, 3, 0)
    a = sampler_oracle.NoiseTape(seed=3).draw_like(x)
    torch.manual_seed(3)
    b = torch.randn_like(x)
    assert torch.equal(a, b) and a.stride() == x.stride()


# ---- hooked branches: inpainting, cond_fn, denoised_fn, BEAT ancestral (tests/golden/make_golden_hooks.py) ----------
sys_path_golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _hook_cases(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_hooks_cases", os.path.join(sys_path_golden, "hook_cases.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, (mod.CASES_TED if name == "ted" else mod.CASES_BEAT)


def run_oracle_hooks(name, tag, dims, sd, B=2):
    mod, cases = _hook_cases(name)
    spec, ddim, seed, kw = cases[tag]
    tab, tmap = schedule_oracle.build("cosine", 1000, spec)
    _, _, _, hooks = mod.hook_objects(kw.get("hooks", ()), dims, B, noised=(name == "ted"))
    tape = sampler_oracle.NoiseTape(seed=seed)
    with torch.no_grad(), few_threads():
        out = sampler_oracle.sample_loop(sd, tab, tmap, (B, dims.njoints, dims.nfeats, 34), synthetic.synth_cond(dims, B),
                                         tape, ddim=ddim, eta=kw.get("eta", 0.0),
                                         clip_denoised=kw.get("clip_denoised", False),
                                         skip_timesteps=kw.get("skip_timesteps", 0),
                                         const_noise=kw.get("const_noise", False), hooks=hooks,
                                         const_noise_init=(name == "ted"))
    return out, tape


HOOK_TAGS = [("ted", t) for t in ("inp_anc", "inp_ddim", "cond_anc", "cond_ddim", "dfn_anc_clip", "all_anc")] + \
            [("beat", t) for t in ("anc100", "anc100_constnoise", "inp_anc", "anc1000_tail")]


@pytest.mark.parametrize("name,tag", [("ted", "inp_ddim"), ("ted", "cond_ddim"), ("ted", "all_anc"),
                                      ("beat", "anc100_constnoise"), ("beat", "inp_anc")])
def test_hooked_loops(name, tag, golden_hooks):
    """Oracle vs the reference's own outputs on the hooked branches (gaussian_diffusion.py:314-320, 429-481)."""
    dims = synthetic.dims_for(name)
    sd = synthetic.synth_state_dict(dims, seed=1)
    out, _ = run_oracle_hooks(name, tag, dims, sd)
    _close(out, golden_hooks[name]["loop_" + tag])


def test_rhythm_metric_oracle(golden_metrics):
    """oracle/metrics_oracle.py against the fixture the reference's own lines produced (scripts/test_RAG_ted.py:84-123
    executed by tests/golden/make_golden_metrics.py): angle changes, motion-beat masks, per-clip and total scores."""
    from oracle import metrics_oracle
    g = golden_metrics
    o = metrics_oracle.motion_beats(torch.from_numpy(g["sample"]), g["mean_dir_vec"], g["angle_pair"], g["change_angle"],
                                    float(g["thres"]))
    np.testing.assert_allclose(o["angle_diff"].numpy(), g["angle_diff"], rtol=0, atol=1e-6)
    assert (o["beat_mask"].numpy() == g["beat_mask"]).all() and int(g["beat_mask"].sum()) == int(g["total_motion"])
    beats = [list(g["audio_beats"][b, :g["audio_n"][b]].astype(np.float64)) for b in range(g["sample"].shape[0])]
    s = metrics_oracle.beat_align(o["beat_mask"], beats, float(g["sigma"]))
    np.testing.assert_allclose(s["clip_score"], g["clip_score"], rtol=1e-6, atol=1e-7)      # onsets stored as fp32
    assert s["total_audio"] == int(g["total_audio"]) and s["total_motion"] == int(g["total_motion"])
    from livelyspeaker_b200 import metrics
    assert list(g["mean_dir_vec"]) == metrics.MEAN_DIR_VEC and [tuple(p) for p in g["angle_pair"]] == metrics.ANGLE_PAIR
    assert list(g["change_angle"]) == metrics.CHANGE_ANGLE and float(g["thres"]) == metrics.THRES
    assert float(g["sigma"]) == metrics.SIGMA


def _vlb_cases():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import vlb_cases
    return vlb_cases


@pytest.mark.parametrize("tag", ["t19", "t1", "t0_near", "loop"])
def test_oracle_variational_bound_vs_reference_fixture(tag, golden_vlb):
    """The oracle's _vb_terms_bpd / _prior_bpd / calc_bpd_loop restatement against the reference's outputs
    (tests/golden/make_golden_vlb.py), TED: KL terms, the discretised decoder NLL away from its clamp, the whole loop."""
    vc = _vlb_cases()
    g = golden_vlb["ted"]
    dims = synthetic.TED
    sd = synthetic.synth_state_dict(dims, seed=1)
    tab, tmap = schedule_oracle.build("cosine", 1000, vc.SPEC)
    shape = (vc.B, dims.njoints, dims.nfeats, 34)
    y = synthetic.synth_cond(dims, vc.B)
    with torch.no_grad():
        if tag == "loop":
            x_start = vc.loop_input(shape)
            o = sampler_oracle.calc_bpd_loop(sd, tab, tmap, x_start, y, sampler_oracle.NoiseTape(seed=vc.LOOP_SEED),
                                             dims.njoints, dims.nfeats, clip_denoised=True)
            for k in ("total_bpd", "prior_bpd", "vb", "xstart_mse", "mse"):
                np.testing.assert_allclose(o[k].numpy(), g["loop_" + k], rtol=2e-5, atol=1e-6)
            np.testing.assert_allclose(sampler_oracle.prior_bpd(tab, x_start).numpy(), g["prior"], rtol=1e-5, atol=1e-8)
            return
        i, seed, clip, near = vc.TERMS[tag]
        x_start, x_t, _ = vc.term_inputs(sampler_oracle.q_sample, tab, shape, i, seed)
        if near:
            x_start = torch.from_numpy(g["x_start_" + tag])
        o, _ = sampler_oracle.vb_terms_bpd(sd, tab, tmap, x_start, x_t, i, y, sampler_oracle.NoiseTape(seed=seed),
                                           dims.njoints, dims.nfeats, clip_denoised=clip)
    np.testing.assert_allclose(o.numpy(), g["vb_" + tag], rtol=2e-5, atol=1e-6)


def _fgd_cases():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import fgd_cases
    return fgd_cases


def test_fgd_oracle_vs_reference_fixture(golden_fgd):
    """oracle/evaluator_oracle.py against the fixture the reference's EmbeddingNet + EmbeddingSpaceEvaluator produced
    (tests/golden/make_golden_fgd.py): features of every pushed batch, Frechet distance, feature distance, diversity."""
    from oracle import evaluator_oracle
    fc = _fgd_cases()
    sd = synthetic.synth_embed_state_dict(seed=fc.SEED_WEIGHTS, pose_dim=fc.POSE_DIM)
    assert abs(sum(float(v.double().abs().sum()) for v in sd.values()) - float(golden_fgd["weights_abs_sum"])) < 1e-6
    gen, real = [], []
    for i, (generated, real_poses) in enumerate(fc.pose_batches()):
        for tag, poses, keep in (("gen", generated, gen), ("real", real_poses, real)):
            mu, logvar = evaluator_oracle.pose_features(sd, poses)
            np.testing.assert_allclose(mu.numpy(), golden_fgd["mu_%s_%d" % (tag, i)], rtol=1e-5, atol=5e-6)
            np.testing.assert_allclose(logvar.numpy(), golden_fgd["logvar_%s_%d" % (tag, i)], rtol=1e-5, atol=5e-6)
            keep.append(golden_fgd["mu_%s_%d" % (tag, i)])
    frechet, feat_dist = evaluator_oracle.scores(gen, real)
    assert abs(frechet - float(golden_fgd["frechet"])) < 1e-9 and abs(feat_dist - float(golden_fgd["feat_dist"])) < 1e-6
    torch.manual_seed(fc.DIVERSITY_SEED)
    assert abs(evaluator_oracle.diversity(gen) - float(golden_fgd["diversity"])) < 1e-6


def test_fgd_host_statistics_and_surface(golden_fgd, tmp_path):
    """The product's EmbeddingSpaceEvaluator: checkpoint contract of the constructor (ted_evaluator.py:14-24), the
    reference's key set (the synthetic gen_dict loaded strictly into the reference module when the fixture was made),
    host-side statistics on the fixture's features, closed-form Frechet cases, and no CPU path for the encoder."""
    from livelyspeaker_b200 import ted_evaluator, embedding_net
    from livelyspeaker_b200._cabi import LsError
    fc = _fgd_cases()
    sd = synthetic.synth_embed_state_dict(seed=fc.SEED_WEIGHTS, pose_dim=fc.POSE_DIM)
    net = embedding_net.EmbeddingNet(fc.POSE_DIM, 34)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    path = str(tmp_path / "gesture_autoencoder_checkpoint_best.bin")
    torch.save({"pose_dim": fc.POSE_DIM, "gen_dict": sd}, path)
    ev = ted_evaluator.EmbeddingSpaceEvaluator(path)
    assert ev.pose_dim == fc.POSE_DIM and not ev.net.training and not any(p.requires_grad for p in ev.net.parameters())
    assert float(ev.net.pose_encoder.out_net[2].negative_slope) == 1.0        # nn.LeakyReLU(True), embedding_net.py:54
    generated, real_poses = fc.pose_batches()[0]
    with pytest.raises(LsError):
        ev.push_samples(generated, real_poses)                               # host tensors: no CPU implementation
    with pytest.raises(NotImplementedError):
        ev.net.train()(generated)
    with pytest.raises(NotImplementedError):
        ev.net.decoder(torch.zeros(1, 32))
    for i in range(fc.N_BATCHES):
        ev.generated_feat_list.append(golden_fgd["mu_gen_%d" % i])
        ev.real_feat_list.append(golden_fgd["mu_real_%d" % i])
    frechet, feat_dist = ev.get_scores()
    assert abs(frechet - float(golden_fgd["frechet"])) < 1e-9 and abs(feat_dist - float(golden_fgd["feat_dist"])) < 1e-6
    torch.manual_seed(fc.DIVERSITY_SEED)
    assert abs(ev.get_diversity_scores() - float(golden_fgd["diversity"])) < 1e-6
    assert ev.get_no_of_samples() == fc.N_BATCHES
    ev.reset()
    assert ev.get_no_of_samples() == 0 and ev.generated_feat_list == []
    # closed forms: identical Gaussians -> 0; commuting diagonal covariances -> |dmu|^2 + sum (sqrt a - sqrt b)^2
    f = ted_evaluator.EmbeddingSpaceEvaluator.calculate_frechet_distance
    a, b = np.array([1.0, 4.0, 9.0]), np.array([4.0, 1.0, 16.0])
    assert abs(f(np.zeros(3), np.diag(a), np.zeros(3), np.diag(a))) < 1e-9
    want = 3 * 0.25 + float(np.sum((np.sqrt(a) - np.sqrt(b)) ** 2))
    assert abs(f(np.full(3, 0.5), np.diag(a), np.zeros(3), np.diag(b)) - want) < 1e-9
    with pytest.raises(AssertionError):
        f(np.zeros(3), np.eye(3), np.zeros(2), np.eye(2))


@pytest.mark.parametrize("tag", ["anc_lo_clip", "ddim_eta_lo"])
def test_oracle_with_grad_samplers_vs_reference_fixture(tag):
    """The oracle's p_sample_with_grad / ddim_sample_with_grad restatement (autograd through the oracle's denoiser)
    against the reference's outputs (tests/golden/make_golden_grad.py), TED."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import grad_cases
    gold = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grad_ted.npz")))
    spec, ddim, seed, i0, n, kw = grad_cases.CASES[tag]
    dims = synthetic.TED
    sd = synthetic.synth_state_dict(dims, seed=1)
    tab, tmap = schedule_oracle.build("cosine", 1000, spec)
    cond_fn, _ = grad_cases.make_cond_fn(dims, 2)
    tape = sampler_oracle.NoiseTape(seed=seed)
    y = synthetic.synth_cond(dims, 2)
    x = tape.draw(2, dims.njoints, dims.nfeats, 34)
    for k in range(n):
        step = sampler_oracle.ddim_sample_with_grad_step if ddim else sampler_oracle.p_sample_with_grad_step
        extra = {"eta": kw.get("eta", 0.0)} if ddim else {}
        x, x0 = step(sd, tab, tmap, x, i0 - k, y, tape, dims.njoints, dims.nfeats, cond_fn,
                     clip_denoised=kw.get("clip_denoised", False), **extra)
        np.testing.assert_allclose(x.numpy(), gold["chain_%s_samples" % tag][k], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(x0.numpy(), gold["chain_%s_x0" % tag][k], rtol=RTOL, atol=ATOL)
