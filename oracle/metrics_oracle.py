"""Oracle (test infrastructure): the post-sampling rhythm metric of the TED evaluation.

torch-CPU restatement of scripts/test_RAG_ted.py:84-123 (infer_from_testloader, between the sampler call and the score
accumulation); pinned by tests/golden/make_golden_metrics.py, which executes those reference lines themselves.

  :84      aligned_motions = sample.permute(0,3,1,2).reshape(B, 34, -1)            [B,F,J*3]
  :88-91   beat_vec = normalize(aligned + mean_dir_vec, per joint)                  direction vectors
  :93-103  per joint pair: angle = acos(clamp(<v1,v2>, -1, 1)) / pi ; angle_diff[t] = sum_pairs |angle[t]-angle[t-1]|
           / change_angle[pair] / n_pairs ; angle_diff[:, 0] = 0
  :106-111 motion beat at frame t in [2, 33): strict local minimum of angle_diff whose drop from either neighbour is
           >= thres ; time = t / 15
  :112-123 per clip: sum over audio onsets of exp(-min_t (onset - beat_t)^2 / (2 sigma^2)) (skipped when the clip has no
           motion beat; such clips do not count their audio onsets either); totals over the batch
The onset detector (librosa.onset.onset_detect, :113) is an input here: parity unpinned, librosa is not available offline.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def motion_beats(sample, mean_dir_vec, angle_pair, change_angle, thres):
    """sample [B,J,3,F] -> {'angle_diff' [B,F] fp32, 'beat_mask' [B,F] uint8}."""
    B, J, D, Fr = sample.shape
    aligned = sample.permute(0, 3, 1, 2).reshape(B, Fr, -1)
    beat_vec = aligned + torch.Tensor(np.asarray(mean_dir_vec)).squeeze()
    beat_vec = F.normalize(beat_vec.reshape(B, Fr, -1, 3), dim=-1)
    all_vec = beat_vec.reshape(B * Fr, -1, 3)
    angle_diff = None
    n = len(change_angle)
    for idx, pair in enumerate(angle_pair):
        v1, v2 = all_vec[:, int(pair[0])], all_vec[:, int(pair[1])]
        inner = torch.clamp(torch.einsum('ij,ij->i', [v1, v2]), -1, 1)
        angle_time = (torch.acos(inner) / math.pi).reshape(B, -1)
        d = torch.abs(angle_time[:, 1:] - angle_time[:, :-1]) / float(change_angle[idx]) / n
        angle_diff = d if idx == 0 else angle_diff + d
    angle_diff = torch.cat((torch.zeros(B, 1), angle_diff), dim=-1)
    mask = torch.zeros(B, Fr, dtype=torch.uint8)
    for b in range(B):
        for t in range(2, Fr - 1):
            a, lo, hi = angle_diff[b][t], angle_diff[b][t - 1], angle_diff[b][t + 1]
            if a < lo and a < hi and (lo - a >= thres or hi - a >= thres):
                mask[b, t] = 1
    return {"angle_diff": angle_diff, "beat_mask": mask}


def beat_align(beat_mask, audio_beats, sigma, fps=15.0):
    """beat_mask [B,F]; audio_beats: per clip list of onset times (s).  Per clip and total scores like :112-123."""
    clip_score, n_audio, n_motion = [], [], []
    for b in range(beat_mask.shape[0]):
        mt = np.asarray([float(t) / fps for t in torch.nonzero(beat_mask[b]).flatten().tolist()])
        n_motion.append(len(mt))
        if len(mt) == 0:
            clip_score.append(0.0)
            n_audio.append(0)
            continue
        s = 0
        for ab in audio_beats[b]:
            s += np.power(math.e, -np.min(np.power((ab - mt), 2)) / (2 * sigma * sigma))
        clip_score.append(float(s))
        n_audio.append(len(audio_beats[b]))
    return {"clip_score": clip_score, "clip_n_audio": n_audio, "clip_n_motion": n_motion,
            "total_score": float(sum(clip_score)), "total_audio": int(sum(n_audio)), "total_motion": int(sum(n_motion))}
