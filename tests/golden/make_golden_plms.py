#!/usr/bin/env python
"""Golden fixtures for the PLMS sampler by RUNNING THE REFERENCE
(/root/reference/scripts/diffusion/gaussian_diffusion.py:1016-1211) in the build container, next to the oracle's
restatement (oracle/sampler_oracle.py::plms_loop) on the same seed: every torch.randn / randn_like draw of the
reference must equal the oracle's tape draw by draw, and the outputs are written to tests/golden/plms.npz.

    python tests/golden/make_golden_plms.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True

from livelyspeaker_b200 import synthetic              # noqa: E402
from oracle import sampler_oracle, schedule_oracle     # noqa: E402
import make_golden as mg                               # noqa: E402

CASES = {   # tag: (respacing, order, seed, kwargs)
    "ddim100_o2": ("ddim100", 2, 301, {}),
    "ddim100_o3": ("ddim100", 3, 302, {}),
    "ddim50_o4_clip": ("ddim50", 4, 303, {"clip_denoised": True}),
    "ddim100_o2_sdedit": ("ddim100", 2, 304, {"skip_timesteps": 80, "init": True}),
}


def main():
    dims = synthetic.TED
    create, load, CFG = mg.import_reference("ted")
    sd = synthetic.synth_state_dict(dims, seed=1)
    out, report = {}, {}
    init = 0.3 * torch.randn(2, dims.njoints, dims.nfeats, synthetic.N_FRAMES, generator=torch.Generator().manual_seed(5))
    out["init_image"] = init.numpy()
    for tag, (spec, order, seed, kw) in CASES.items():
        model, diffusion = create(mg.ref_args(dims), spec)
        load(model, {k: v.clone() for k, v in sd.items()})
        cfg = CFG(model).eval()
        tab, tmap = schedule_oracle.build("cosine", 1000, spec)
        shp = (2, dims.njoints, dims.nfeats, synthetic.N_FRAMES)
        ii = init if kw.get("init") else None
        drawn = []
        y_ref = mg.fresh_cond(dims, 2)          # before the draw recorder is installed (synth_cond uses torch.randn)
        o_randn, o_like = torch.randn, torch.randn_like
        torch.randn = lambda *a_, **k_: (drawn.append(o_randn(*a_, **k_)), drawn[-1])[1]
        torch.randn_like = lambda *a_, **k_: (drawn.append(o_like(*a_, **k_)), drawn[-1])[1]
        try:
            torch.manual_seed(seed)
            with torch.no_grad():
                r = diffusion.plms_sample_loop(cfg, shp, clip_denoised=kw.get("clip_denoised", False),
                                               model_kwargs={"y": y_ref},
                                               skip_timesteps=kw.get("skip_timesteps", 0), init_image=ii, order=order)
        finally:
            torch.randn, torch.randn_like = o_randn, o_like
        tape = sampler_oracle.NoiseTape(seed=seed)
        with torch.no_grad():
            o = sampler_oracle.plms_loop(sd, tab, tmap, shp, mg.fresh_cond(dims, 2), tape, order=order,
                                         clip_denoised=kw.get("clip_denoised", False),
                                         skip_timesteps=kw.get("skip_timesteps", 0), init_image=ii)
        assert len(drawn) == len(tape.record), (tag, len(drawn), len(tape.record))
        assert all(torch.equal(p_, q_) for p_, q_ in zip(drawn, tape.record)), tag
        out["plms_" + tag] = r.numpy()
        report["plms_%s" % tag] = mg.maxabs(r, o)
        report["plms_%s_draws" % tag] = len(drawn)
        print(tag, "draws", len(drawn), "max|ref-oracle|", report["plms_%s" % tag], flush=True)
    np.savez_compressed(os.path.join(HERE, "plms.npz"), **out)
    rp = os.path.join(HERE, "PIN_REPORT.json")
    allrep = json.load(open(rp)) if os.path.exists(rp) else {}
    allrep["plms"] = report
    json.dump(allrep, open(rp, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
