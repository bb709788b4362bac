#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Runs only in the build container (needs /root/reference); the GPU box and the
test-suite never import the reference - they read the .npz files this wrote.
While generating, every fixture is also replayed through ``oracle/`` and the
max-abs differences are written to PIN_REPORT.json: that is the oracle's pin.

    python tests/golden/make_golden.py --dataset ted
    python tests/golden/make_golden.py --dataset beat

One reference tree per process (scripts/ and scripts_beat/ both define the
top-level packages `model`, `diffusion`, `mdm_utils`).  Import recipe from
SURVEY.md section 8c: stub `clip` (imported, never used: scripts/model/RAG.py:5) and
disable bytecode writes (the reference mount is read-only).
"""
import argparse
import json
import os
import sys
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from livelyspeaker_b200 import synthetic
from oracle import rag_oracle, sampler_oracle, schedule_oracle

TABLE_KEYS = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next",
              "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod",
              "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
              "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]

SCHEDULE_CASES = [("cosine", 1000, ""), ("cosine", 1000, "ddim100"), ("cosine", 1000, "ddim50"),
                  ("cosine", 1000, "250"), ("cosine", 1000, "100"), ("cosine", 300, "10,15,20"),
                  ("linear", 1000, "ddim20"), ("cosine", 100, ""), ("cosine", 1000, "7,3")]


def import_reference(dataset):
    tree = "/root/reference/scripts" if dataset == "ted" else "/root/reference/scripts_beat"
    sys.path.insert(0, tree)
    sys.modules["clip"] = types.ModuleType("clip")
    import warnings
    warnings.filterwarnings("ignore")
    from mdm_utils.model_util import create_model_and_diffusion, load_model_wo_clip
    from model.cfg_sampler import ClassifierFreeSampleModel
    return create_model_and_diffusion, load_model_wo_clip, ClassifierFreeSampleModel


def ref_args(dims, diffusion_steps=1000, noise_schedule="cosine"):
    return types.SimpleNamespace(
        mdm_condm="text", latent_dim=dims.latent_dim, ff_size=1024, layers=dims.layers,
        cond_mask_prob=0.1, arch="trans_enc", emb_trans_dec=False, dataset="humanml",
        lang_model=None, mlpact="silu", diffusion_steps=diffusion_steps,
        noise_schedule=noise_schedule, sigma_small=True, lambda_vel=1.0, lambda_rcxyz=0.0,
        lambda_fc=0.0, njoints=dims.njoints)


def fresh_cond(dims, B, seed=233):
    return synthetic.synth_cond(dims, B, seed=seed)


def maxabs(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="ted", choices=["ted", "beat"])
    a = ap.parse_args()
    dims = synthetic.dims_for(a.dataset)
    create, load_wo_clip, CFG = import_reference(a.dataset)
    torch.set_num_threads(8)
    sd = synthetic.synth_state_dict(dims, seed=1)
    report = {}
    out = {}

    # ---- schedule / respacing (bit-exact fp64) ------------------------------------
    if a.dataset == "ted":
        sched = {}
        for k, (name, steps, spec) in enumerate(SCHEDULE_CASES):
            _, diff = create(ref_args(dims, steps, name), spec)
            tab, tmap = schedule_oracle.build(name, steps, spec)
            assert list(diff.timestep_map) == list(tmap), (name, steps, spec)
            for key in TABLE_KEYS:
                ref = getattr(diff, key)
                assert np.array_equal(ref, tab[key]), (name, steps, spec, key)
                sched["c%d_%s" % (k, key)] = ref
            sched["c%d_timestep_map" % k] = np.array(diff.timestep_map, dtype=np.int64)
            sched["c%d_num_timesteps" % k] = np.array(diff.num_timesteps)
        sched["cases"] = np.array(json.dumps(SCHEDULE_CASES))
        np.savez_compressed(os.path.join(HERE, "schedule.npz"), **sched)
        report["schedule"] = "bit-exact on %d cases x %d tables + timestep_map" % (len(SCHEDULE_CASES), len(TABLE_KEYS))

    # ---- model ----------------------------------------------------------------------
    model, diff_ddim = create(ref_args(dims), "ddim100")
    load_wo_clip(model, {k: v.clone() for k, v in sd.items()})
    model.eval()
    cfg = CFG(model).eval()
    B = 2
    nj, nf = dims.njoints, dims.nfeats
    shape = (B, nj, nf, synthetic.N_FRAMES)
    out["weights_abs_sum"] = np.array(sum(float(v.double().abs().sum()) for v in sd.values()))

    with torch.no_grad():
        # per-op goldens (TED only; the BEAT twin shares the op code)
        g = torch.Generator().manual_seed(7)
        if a.dataset == "ted":
            y = fresh_cond(dims, B)
            af = model.audio_encoder(y["audio_input"], num_frames=34)
            out["wavenc_out"] = af.numpy()
            report["wavenc"] = maxabs(af, rag_oracle.wav_encoder(sd, y["audio_input"]))
            hx = torch.randn(B, dims.seq_len, dims.latent_dim, generator=g) * 2 + 0.5
            out["block_in"] = hx.numpy()
            ln = model.backbone.mlps[3].block1[0](hx)
            out["ln_out"] = ln.numpy()
            report["ln"] = maxabs(ln, rag_oracle.ln_spatial(hx, sd["backbone.mlps.3.block1.0.alpha"],
                                                            sd["backbone.mlps.3.block1.0.beta"]))
            tt = torch.tensor([990, 17])
            emb = model.backbone.embed_timestep(tt)
            out["temb_t"] = tt.numpy()
            out["temb_out"] = emb.numpy()
            report["temb"] = maxabs(emb, rag_oracle.timestep_embed(sd, tt))
            blk = model.backbone.mlps[3](hx, emb)
            out["block_out"] = blk.numpy()
            report["block"] = maxabs(blk, rag_oracle.mlp_block(sd, 3, hx, emb))
            bb = model.backbone(hx, tt)
            out["backbone_out"] = bb.numpy()
            report["backbone"] = maxabs(bb, rag_oracle.trans_mlp(sd, hx, tt))

        # RAG.forward cond / uncond, CFG
        x = torch.randn(*shape, generator=g)
        t = torch.tensor([730, 40])
        out["fwd_x"] = x.numpy()
        out["fwd_t"] = t.numpy()
        for tag, unc in (("cond", False), ("uncond", True)):
            y = fresh_cond(dims, B)
            if unc:
                y["uncond"] = True
            torch.manual_seed(11)
            r = model(x, t, y)
            eps = torch.randn(B, 1, dims.latent_dim, generator=torch.Generator().manual_seed(11))
            y2 = fresh_cond(dims, B)
            if unc:
                y2["uncond"] = True
            o = rag_oracle.rag_forward(sd, x, t, y2, eps, nj, nf)
            for key in ("output", "z_mu", "z_logvar"):
                out["fwd_%s_%s" % (tag, key)] = r[key].numpy()
                report["fwd_%s_%s" % (tag, key)] = maxabs(r[key], o[key])
            assert float(y["origin_x"][..., 4:].abs().max()) == 0.0   # in-place side effect
        y = fresh_cond(dims, B)
        torch.manual_seed(12)
        c = cfg(x, t, y)
        tape = sampler_oracle.NoiseTape(seed=12)
        o = rag_oracle.cfg_forward(sd, x, t, fresh_cond(dims, B), tape.draw(B, 1, 512), tape.draw(B, 1, 512), nj, nf)
        out["cfg_out"] = c.numpy()
        report["cfg"] = maxabs(c, o)

        # single steps (ancestral on the full 1000-step process, DDIM on ddim100)
        _, diff_full = create(ref_args(dims), "")
        tab_full, map_full = schedule_oracle.build("cosine", 1000, "")
        tab_ddim, map_ddim = schedule_oracle.build("cosine", 1000, "ddim100")
        xs = torch.randn(*shape, generator=g)
        out["step_x"] = xs.numpy()
        for i in (999, 500, 1, 0):
            torch.manual_seed(100 + i)
            r = diff_full.p_sample(cfg, xs, torch.tensor([i] * B), clip_denoised=False,
                                   model_kwargs={"y": fresh_cond(dims, B)})
            tape = sampler_oracle.NoiseTape(seed=100 + i)
            s, p0 = sampler_oracle.p_sample_step(sd, tab_full, map_full, xs, i, fresh_cond(dims, B), tape, nj, nf)
            out["pstep%d_sample" % i] = r["sample"].numpy()
            out["pstep%d_x0" % i] = r["pred_xstart"].numpy()
            report["pstep%d" % i] = max(maxabs(r["sample"], s), maxabs(r["pred_xstart"], p0))
        for i, eta in ((99, 0.0), (57, 0.0), (57, 0.5), (0, 0.0)):
            torch.manual_seed(200 + i)
            r = diff_ddim.ddim_sample(cfg, xs, torch.tensor([i] * B), clip_denoised=False,
                                      model_kwargs={"y": fresh_cond(dims, B)}, eta=eta)
            tape = sampler_oracle.NoiseTape(seed=200 + i)
            s, p0 = sampler_oracle.ddim_step(sd, tab_ddim, map_ddim, xs, i, fresh_cond(dims, B), tape, nj, nf, eta=eta)
            tag = "dstep%d_eta%d" % (i, int(eta * 10))
            out[tag + "_sample"] = r["sample"].numpy()
            out[tag + "_x0"] = r["pred_xstart"].numpy()
            report[tag] = max(maxabs(r["sample"], s), maxabs(r["pred_xstart"], p0))

        # whole loops
        def run_loop(tag, diffusion, tab, tmap, ddim, seed, Bn=2, **kw):
            shp = (Bn, nj, nf, synthetic.N_FRAMES)
            torch.manual_seed(seed)
            fn = diffusion.ddim_sample_loop if ddim else diffusion.p_sample_loop
            init = kw.get("init_image")
            drawn = []
            y_ref = fresh_cond(dims, Bn)
            o_randn, o_like = torch.randn, torch.randn_like
            torch.randn = lambda *a_, **k_: (drawn.append(o_randn(*a_, **k_)), drawn[-1])[1]
            torch.randn_like = lambda *a_, **k_: (drawn.append(o_like(*a_, **k_)), drawn[-1])[1]
            try:
                r = self_call(fn, shp, y_ref, init, kw)
            finally:
                torch.randn, torch.randn_like = o_randn, o_like
            tape = sampler_oracle.NoiseTape(seed=seed)
            o = sampler_oracle.sample_loop(sd, tab, tmap, shp, fresh_cond(dims, Bn), tape, ddim=ddim,
                                           eta=kw.get("eta", 0.0), clip_denoised=kw.get("clip_denoised", False),
                                           skip_timesteps=kw.get("skip_timesteps", 0), init_image=init,
                                           const_noise=kw.get("const_noise", False))
            # RNG-order pin: the oracle's tape must equal the reference's own draws, one by one
            assert len(drawn) == len(tape.record), (len(drawn), len(tape.record))
            assert all(torch.equal(p_, q_) for p_, q_ in zip(drawn, tape.record)), tag
            report["loop_%s_draws" % tag] = len(drawn)
            out["loop_%s" % tag] = r.numpy()
            report["loop_%s" % tag] = maxabs(r, o)
            print("loop", tag, report["loop_%s" % tag], flush=True)

        def self_call(fn, shp, y_ref, init, kw):
            return fn(cfg, shp, clip_denoised=kw.get("clip_denoised", False),
                   model_kwargs={"y": y_ref}, skip_timesteps=kw.get("skip_timesteps", 0),
                   init_image=init, progress=False, dump_steps=None, noise=None,
                   const_noise=kw.get("const_noise", False), **({"eta": kw["eta"]} if "eta" in kw else {}))

        run_loop("ddim100", diff_ddim, tab_ddim, map_ddim, True, 233)
        if a.dataset == "ted":
            _, diff_100 = create(ref_args(dims), "100")
            tab_100, map_100 = schedule_oracle.build("cosine", 1000, "100")
            run_loop("anc100", diff_100, tab_100, map_100, False, 234)
            init = 0.3 * torch.randn(*shape, generator=torch.Generator().manual_seed(5))
            out["init_image"] = init.numpy()
            run_loop("ddim100_sdedit", diff_ddim, tab_ddim, map_ddim, True, 235, init_image=init, skip_timesteps=80)
            run_loop("ddim100_eta05_clip", diff_ddim, tab_ddim, map_ddim, True, 236, eta=0.5, clip_denoised=True)
            run_loop("anc100_constnoise", diff_100, tab_100, map_100, False, 237, const_noise=True)
            run_loop("anc100_skip60", diff_100, tab_100, map_100, False, 238, skip_timesteps=60)
            # config 1 parity gate: 1 clip, T=100 (diffusion_steps=100), ancestral
            _, diff_t100 = create(ref_args(dims, 100), "")
            tab_t100, map_t100 = schedule_oracle.build("cosine", 100, "")
            run_loop("t100_b1", diff_t100, tab_t100, map_t100, False, 239, Bn=1)
            # full T=1000 ancestral, 1 clip
            run_loop("anc1000_b1", diff_full, tab_full, map_full, False, 240, Bn=1)

    np.savez_compressed(os.path.join(HERE, "%s.npz" % a.dataset), **out)
    rp = os.path.join(HERE, "PIN_REPORT.json")
    allrep = json.load(open(rp)) if os.path.exists(rp) else {}
    allrep[a.dataset] = report
    allrep["_how"] = ("max-abs |reference - oracle| per fixture, reference = /root/reference @ 7f6ccd1 "
                      "on torch %s CPU fp32" % torch.__version__)
    json.dump(allrep, open(rp, "w"), indent=1, sort_keys=True)
    worst = max(v for v in report.values() if isinstance(v, float))
    print("worst |ref-oracle| =", worst)
    assert worst < 2e-5, report


if __name__ == "__main__":
    main()
