"""TEST INFRASTRUCTURE - CPU restatement of the feature side of the TED evaluation:

  scripts/model/embedding_net.py:15-37   ConvNormRelu (Conv1d k3/s1 or k4/s2, BatchNorm1d, LeakyReLU(0.2))
  scripts/model/embedding_net.py:40-79   PoseEncoderConv: four convolutions, flatten, three linear layers, fc_mu / fc_logvar
  scripts/model/ted_evaluator.py:35-41   push_samples: the mu head of generated and real clips
  scripts/model/ted_evaluator.py:59-88   get_scores: Frechet distance of the two feature clouds + mean L1 feature distance
  scripts/model/ted_evaluator.py:90-145  calculate_frechet_distance (scipy.linalg.sqrtm of the covariance product)
  scripts/model/ted_evaluator.py:147-154 get_diversity_scores

BatchNorm is evaluated on its running statistics (the evaluator puts the net in eval mode) and written out as plain
arithmetic; the linear stack's nn.LeakyReLU(True) has negative_slope = True = 1.0, i.e. it is the identity.  Pinned
against the reference modules by tests/golden/make_golden_fgd.py (tests/golden/fgd.npz).  Only tests/, smoke() and
bench.py may import this.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy import linalg

BN_EPS = 1e-5
SLOPE_CONV = 0.2
SLOPE_FC = 1.0           # nn.LeakyReLU(True)


def _bn(sd, p, x):
    """Eval-mode BatchNorm1d over dim 1 of [B, C] or [B, C, L]."""
    shape = (1, -1) + (1,) * (x.dim() - 2)
    inv = torch.rsqrt(sd[p + "running_var"] + BN_EPS)
    return (x - sd[p + "running_mean"].view(shape)) * (inv * sd[p + "weight"]).view(shape) + sd[p + "bias"].view(shape)


def pose_features(sd, poses, prefix="pose_encoder."):
    """poses [B, 34, dim] -> (mu [B, 32], logvar [B, 32]); embedding_net.py:64-79 with variational_encoding=False."""
    p = prefix
    x = poses.float().transpose(1, 2)
    for i, stride in ((0, 1), (1, 1), (2, 2)):
        x = F.conv1d(x, sd[p + "net.%d.0.weight" % i], sd[p + "net.%d.0.bias" % i], stride=stride)
        x = F.leaky_relu(_bn(sd, p + "net.%d.1." % i, x), SLOPE_CONV)
    x = F.conv1d(x, sd[p + "net.3.weight"], sd[p + "net.3.bias"])
    x = x.flatten(1)
    for lin, bn in ((0, 1), (3, 4)):
        x = x @ sd[p + "out_net.%d.weight" % lin].t() + sd[p + "out_net.%d.bias" % lin]
        x = F.leaky_relu(_bn(sd, p + "out_net.%d." % bn, x), SLOPE_FC)
    x = x @ sd[p + "out_net.6.weight"].t() + sd[p + "out_net.6.bias"]
    mu = x @ sd[p + "fc_mu.weight"].t() + sd[p + "fc_mu.bias"]
    logvar = x @ sd[p + "fc_logvar.weight"].t() + sd[p + "fc_logvar.bias"]
    return mu, logvar


def _sqrtm(m):
    """scipy.linalg.sqrtm; releases before 1.16 print a warning unless told disp=False (what the reference passes)."""
    try:
        return linalg.sqrtm(m, disp=False)[0]
    except TypeError:
        return linalg.sqrtm(m)


def frechet_distance(feats_a, feats_b, eps=1e-6):
    """ted_evaluator.py:62-71, 90-145: d^2 = |mu_a - mu_b|^2 + tr(S_a + S_b - 2 (S_a S_b)^(1/2)) of Gaussian fits."""
    mu_a, mu_b = feats_a.mean(axis=0), feats_b.mean(axis=0)
    s_a, s_b = np.cov(feats_a, rowvar=False), np.cov(feats_b, rowvar=False)
    root = _sqrtm(s_a.dot(s_b))
    if not np.isfinite(root).all():
        jitter = np.eye(s_a.shape[0]) * eps
        root = _sqrtm((s_a + jitter).dot(s_b + jitter))
    if np.iscomplexobj(root):
        if not np.allclose(np.diagonal(root).imag, 0, atol=1e-3):
            return float("inf")              # the reference's ValueError -> 1e+10000000000000 (:69-70)
        root = root.real
    d = mu_a - mu_b
    return d.dot(d) + np.trace(s_a) + np.trace(s_b) - 2 * np.trace(root)


def scores(generated_feat_list, real_feat_list):
    """get_scores (:59-88): (frechet_dist, feat_dist)."""
    gen, real = np.vstack(generated_feat_list), np.vstack(real_feat_list)
    return frechet_distance(gen, real), float(np.mean(np.sum(np.abs(real - gen), axis=1)))


def diversity(generated_feat_list):
    """get_diversity_scores (:147-154); draws one torch.randperm from the global generator."""
    first = np.vstack(generated_feat_list[:500])
    order = torch.randperm(len(generated_feat_list))[:500]
    other = np.vstack([generated_feat_list[i] for i in order])
    return np.mean(np.sum(np.absolute(first - other), axis=-1))
