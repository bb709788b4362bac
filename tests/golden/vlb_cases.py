"""Cases of the variational-bound fixture (vlb_ted.npz / vlb_beat.npz): shared by the generator (make_golden_vlb.py, runs
the reference) and the tests (never import the reference).  GaussianDiffusion._vb_terms_bpd / _prior_bpd / calc_bpd_loop
(scripts/diffusion/gaussian_diffusion.py:1213-1247, 1573-1645) on a 20-step respacing, two clips."""
SPEC = "ddim20"
B = 2
# tag: (spaced index i, seed, clip_denoised, x_start placed near the model's mean)
TERMS = {
    "t19": (19, 601, True, False),
    "t10_noclip": (10, 602, False, False),
    "t1": (1, 603, True, False),
    "t0_far": (0, 604, True, False),          # decoder NLL, every bin saturated at the 1e-12 clamp
    "t0_near": (0, 605, True, True),          # decoder NLL, x_start within a few sigma of the mean: all three branches
    "t0_near_noclip": (0, 606, False, True),
}
MIXED = ("mixed", (0, 7), 607, True)          # per-clip timesteps in one call (reference output only)
LOOP_SEED = 610


def term_inputs(q_sample, tab, shape, i, seed):
    """x_start, x_t of a _vb_terms_bpd case ('near' cases replace x_start by the stored one); also returns the generator."""
    import torch
    g = torch.Generator().manual_seed(seed + 1000)
    x_start = 0.5 * torch.randn(*shape, generator=g)
    x_t = q_sample(tab, x_start, i, torch.randn(*shape, generator=g))
    return x_start, x_t, g


def mixed_inputs(q_sample, tab, shape):
    import torch
    tag, ts, seed, clip = MIXED
    g = torch.Generator().manual_seed(seed + 1000)
    x_start = 0.5 * torch.randn(*shape, generator=g)
    nz = torch.randn(*shape, generator=g)
    x_t = torch.stack([q_sample(tab, x_start[b], int(ts[b]), nz[b]) for b in range(shape[0])])
    return x_start, x_t


def loop_input(shape):
    import torch
    g = torch.Generator().manual_seed(LOOP_SEED + 1000)
    return 0.5 * torch.randn(*shape, generator=g)
