#!/usr/bin/env python
"""Golden fixture for the FGD / diversity scores by RUNNING THE REFERENCE:
/root/reference/scripts/model/embedding_net.py::EmbeddingNet (eval mode, CPU) with the synthetic `gen_dict` loaded
strictly, driven through the reference's own EmbeddingSpaceEvaluator.push_samples / get_scores /
get_diversity_scores (scripts/model/ted_evaluator.py).  The evaluator's constructor needs the private checkpoint and a
GPU, so the object is made without it and handed the net; `umap` (visualisation only) is stubbed, and because this
image's scipy (1.18) no longer takes the `disp=False` argument the reference passes to `linalg.sqrtm`, the module's
`linalg` is wrapped to accept it and return the old `(root, error_estimate)` pair.  Also checks the oracle's restatement (oracle/evaluator_oracle.py) against it.  Writes tests/golden/fgd.npz.

    python tests/golden/make_golden_fgd.py
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True

from livelyspeaker_b200 import synthetic     # noqa: E402
from oracle import evaluator_oracle          # noqa: E402
import fgd_cases                              # noqa: E402


def main():
    sys.path.insert(0, "/root/reference/scripts")
    sys.modules["umap"] = types.ModuleType("umap")
    from model.embedding_net import EmbeddingNet
    import model.ted_evaluator as ref_eval
    from model.ted_evaluator import EmbeddingSpaceEvaluator
    from scipy import linalg as sp_linalg

    def sqrtm_old_api(a, disp=True, blocksize=None):
        root = sp_linalg.sqrtm(a)
        return root if disp else (root, 0.0)
    ref_eval.linalg = types.SimpleNamespace(sqrtm=sqrtm_old_api)
    sd = synthetic.synth_embed_state_dict(seed=fgd_cases.SEED_WEIGHTS, pose_dim=fgd_cases.POSE_DIM)
    net = EmbeddingNet(fgd_cases.POSE_DIM, 34)
    net.load_state_dict(sd, strict=True)
    net.train(False)
    net.freeze_pose_nets()
    ev = object.__new__(EmbeddingSpaceEvaluator)
    ev.pose_dim, ev.net = fgd_cases.POSE_DIM, net
    ev.reset()
    out, worst = {}, 0.0
    for i, (generated, real) in enumerate(fgd_cases.pose_batches()):
        with torch.no_grad():
            ev.push_samples(generated, real)
            for tag, poses in (("gen", generated), ("real", real)):
                z, mu, logvar = net.pose_encoder(poses, False)
                o_mu, o_lv = evaluator_oracle.pose_features(sd, poses)
                worst = max(worst, float((mu - o_mu).abs().max()), float((logvar - o_lv).abs().max()))
                out["mu_%s_%d" % (tag, i)] = mu.numpy()
                out["logvar_%s_%d" % (tag, i)] = logvar.numpy()
        assert np.array_equal(ev.generated_feat_list[-1], out["mu_gen_%d" % i])
    frechet, feat_dist = ev.get_scores()
    torch.manual_seed(fgd_cases.DIVERSITY_SEED)
    diversity = ev.get_diversity_scores()
    o_frechet, o_feat = evaluator_oracle.scores(ev.generated_feat_list, ev.real_feat_list)
    torch.manual_seed(fgd_cases.DIVERSITY_SEED)
    o_div = evaluator_oracle.diversity(ev.generated_feat_list)
    print("max |reference - oracle| features =", worst)
    print("frechet %.9g (oracle %.9g)  feat_dist %.9g (oracle %.9g)  diversity %.9g (oracle %.9g)  samples %d"
          % (frechet, o_frechet, feat_dist, o_feat, diversity, o_div, ev.get_no_of_samples()))
    assert worst < 2e-5 and abs(frechet - o_frechet) < 1e-9 * max(1.0, abs(frechet))
    assert abs(feat_dist - o_feat) < 1e-6 and abs(diversity - o_div) < 1e-6
    out.update(frechet=np.float64(frechet), feat_dist=np.float64(feat_dist), diversity=np.float64(diversity),
               weights_abs_sum=np.float64(sum(float(v.double().abs().sum()) for v in sd.values())))
    np.savez_compressed(os.path.join(HERE, "fgd.npz"), **out)
    rp = os.path.join(HERE, "PIN_REPORT.json")
    allrep = json.load(open(rp)) if os.path.exists(rp) else {}
    allrep["fgd"] = {"features": worst, "keys": len(sd), "frechet": float(frechet), "feat_dist": float(feat_dist),
                     "diversity": float(diversity)}
    json.dump(allrep, open(rp, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
