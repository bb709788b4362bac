"""Inputs of the FGD fixture (fgd.npz): shared by the generator (make_golden_fgd.py, runs the reference) and the tests
(never import the reference).  Pose batches in the layout push_samples receives (scripts/test_RAG_ted.py:84-86):
[B, 34, 27] mean-subtracted direction vectors, `generated` a perturbed copy of `real`."""
import torch

SEED_WEIGHTS = 5
POSE_DIM = 27
N_BATCHES = 5
BATCH = 24
DIVERSITY_SEED = 11


def pose_batches():
    g = torch.Generator().manual_seed(31)
    out = []
    for i in range(N_BATCHES):
        real = 0.3 * torch.randn(BATCH, 34, POSE_DIM, generator=g)
        generated = real + 0.1 * torch.randn(BATCH, 34, POSE_DIM, generator=g) + 0.02 * (i + 1)
        out.append((generated, real))
    return out
