#!/usr/bin/env python
"""Drop-in usage of the two inference flows of the reference, on synthetic data (no datasets / checkpoints offline).

  python examples/sample_synthetic.py --mode rag     # scripts/test_RAG_ted.py:146-182 + infer_from_testloader :43-82
  python examples/sample_synthetic.py --mode lively  # scripts/test_LivelySpeaker_ted.py:80-113 (SAG decoder -> init_image)

Only the import line differs from the reference scripts: `from livelyspeaker_b200 import ...` instead of
`from mdm_utils.model_util import ...` / `from model.cfg_sampler import ...`.  A real run would `torch.load` the
reference's checkpoints (`ckpts/TED/RAG.pt`, `ckpts/TED/SAG.pth`); here the same key sets are filled with seeded
values and saved / reloaded through the checkpoint format the reference uses (flat fp32 state_dict).
"""
import argparse
import os
import sys
import tempfile
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from livelyspeaker_b200 import (ClassifierFreeSampleModel, Decoder_TRANSFORMER, create_model_and_diffusion,  # noqa: E402
                                load_model_wo_clip, metrics, synthetic)
from livelyspeaker_b200.ted_evaluator import EmbeddingSpaceEvaluator  # noqa: E402  (model.ted_evaluator in the reference)


def generate_args():
    return types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1,
                                 arch='trans_enc', emb_trans_dec=False, dataset='humanml', lang_model=None, mlpact='silu',
                                 diffusion_steps=1000, noise_schedule='cosine', sigma_small=True, lambda_vel=1.0,
                                 lambda_rcxyz=0.0, lambda_fc=0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="rag", choices=["rag", "lively"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ddim", type=int, default=1)
    a = ap.parse_args()
    device = torch.device("cuda:0")
    torch.manual_seed(233)                                            # fixseed(233)
    dims = synthetic.TED
    B = a.batch

    # ---- checkpoint round trip in the reference's format
    tmp = tempfile.mkdtemp()
    torch.save(synthetic.synth_state_dict(dims, seed=1), os.path.join(tmp, "RAG.pt"))
    torch.save(synthetic.synth_sag_state_dict(seed=3), os.path.join(tmp, "SAG.pth"))

    model, diffusion = create_model_and_diffusion(generate_args(), 'ddim100' if a.ddim else '')
    load_model_wo_clip(model, torch.load(os.path.join(tmp, "RAG.pt"), map_location='cpu'))
    model = ClassifierFreeSampleModel(model)                          # wrapping model with the classifier-free sampler
    model.to(device)
    model.eval()                                                      # disable random masking
    sample_fn = diffusion.ddim_sample_loop if a.ddim else diffusion.p_sample_loop

    y = synthetic.synth_cond(dims, B, device=device, scale=1.0)       # stands in for one batch of the TED test loader
    cond = {'y': y}
    init_image, skip_steps = None, 0
    if a.mode == "lively":
        sag = Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False)
        sag.load_state_dict(torch.load(os.path.join(tmp, "SAG.pth"), map_location='cpu'))
        sag = sag.to(device).eval()
        z = torch.randn(B, 512, device=device)                        # clip_model.encode_text(texts) in the reference
        batch = {"x": y['origin_x'].clone(), 'mask': torch.ones(B, 34, device=device).bool(), 'z': z}
        init_image = sag(batch)['output']                             # decoded_motions
        skip_steps = 80 if a.ddim else 800
    for attempt in range(2):       # the first call also builds the engine (weight upload, bf16 tapes, embedding table)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            sample = sample_fn(model, (B, model.njoints, model.nfeats, 34), clip_denoised=False, model_kwargs=cond,
                               skip_timesteps=skip_steps, init_image=init_image, progress=False, dump_steps=None,
                               noise=None, const_noise=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    aligned_motions = sample.permute(0, 3, 1, 2).reshape(B, 34, -1)   # what the evaluators consume
    n = diffusion.num_timesteps - skip_steps
    print("%s: %d clips x %d steps in %.1f ms (%.0f clips/s), motions %s, finite=%s"
          % (a.mode, B, n, dt * 1e3, B / dt, tuple(aligned_motions.shape), bool(torch.isfinite(aligned_motions).all())))

    # ---- the script's metrics on the sampler's output, still on the device (scripts/test_RAG_ted.py:35, 84-131)
    torch.save({"pose_dim": 27, "gen_dict": synthetic.synth_embed_state_dict(seed=5)}, os.path.join(tmp, "autoencoder.bin"))
    embed_space_evaluator = EmbeddingSpaceEvaluator(os.path.join(tmp, "autoencoder.bin"))
    vec_seq = y['origin_x'].permute(0, 3, 1, 2).reshape(B, 34, -1)
    embed_space_evaluator.push_samples(aligned_motions, vec_seq)
    frechet_dist, feat_dist = embed_space_evaluator.get_scores()
    _, beat_mask = metrics.motion_beats(sample)
    onsets = [[0.2 * (k + 1) for k in range(10)] for _ in range(B)]  # librosa.onset.onset_detect(...) in the reference
    s = metrics.beat_align_score(beat_mask, onsets)
    print("metrics (synthetic weights, so the values mean nothing): FGD %.4f, feat_dist %.4f, diversity %.4f, beat_score %.4f,"
          " motion_beats_sum %d" % (frechet_dist, feat_dist, embed_space_evaluator.get_diversity_scores(),
                                    s["beat_align_score_sum"] / max(1, s["num_beats"]), s["motion_beats_sum"]))


if __name__ == "__main__":
    main()
