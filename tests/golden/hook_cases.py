"""Case tables and the deterministic hook objects of the hooked-branch fixtures (hooks_ted.npz / hooks_beat.npz).
Shared by the generator (make_golden_hooks.py, runs the reference) and the tests (never import the reference)."""
import torch

from livelyspeaker_b200 import synthetic

# tag: (respacing, ddim, seed, kwargs).  hooks are named and rebuilt identically by the tests (hook_objects).
CASES_TED = {
    "inp_anc": ("100", False, 401, {"skip_timesteps": 80, "hooks": ("inpaint",)}),
    "inp_ddim": ("ddim100", True, 402, {"skip_timesteps": 85, "hooks": ("inpaint",)}),
    "cond_anc": ("100", False, 403, {"skip_timesteps": 85, "hooks": ("cond_fn",)}),
    "cond_ddim": ("ddim100", True, 404, {"skip_timesteps": 85, "eta": 0.2, "hooks": ("cond_fn",)}),
    "dfn_anc_clip": ("100", False, 405, {"skip_timesteps": 90, "clip_denoised": True, "hooks": ("denoised_fn",)}),
    "all_anc": ("100", False, 406, {"skip_timesteps": 92, "hooks": ("inpaint", "cond_fn", "denoised_fn")}),
}
CASES_BEAT = {
    "anc100": ("100", False, 411, {"skip_timesteps": 80}),
    "anc100_constnoise": ("100", False, 412, {"skip_timesteps": 90, "const_noise": True}),
    "inp_anc": ("100", False, 413, {"skip_timesteps": 90, "hooks": ("inpaint",)}),
    "anc1000_tail": ("", False, 414, {"skip_timesteps": 980}),
}


def hook_objects(names, dims, B, noised):
    """The deterministic hook set of a case: (reference-side model_kwargs additions, cond_fn, denoised_fn, oracle hooks)."""
    g = torch.Generator().manual_seed(77)
    shape = (B, dims.njoints, dims.nfeats, synthetic.N_FRAMES)
    y_extra, cond_fn, denoised_fn, hooks = {}, None, None, {}
    if "inpaint" in names:
        mask = torch.zeros(shape, dtype=torch.bool)
        mask[:, : max(1, dims.njoints // 3), :, :12] = True        # the first joints are known on the first 12 frames
        motion = 0.3 * torch.randn(shape, generator=g)
        y_extra = {"inpainting_mask": mask, "inpainted_motion": motion}
        hooks["inpaint"] = (mask, motion, noised)
    if "cond_fn" in names:
        target = 0.2 * torch.randn(shape, generator=g)

        def cond_fn(x, t, y=None):
            # gradient of -0.5*w(t)*|x - target|^2 ; depends on x, the (ORIGINAL) timestep and the batch index
            w = 0.5 + t.float().view(-1, 1, 1, 1) / 1000.0
            return -w * (x - target.to(x.device))
        hooks["cond_fn"] = cond_fn
    if "denoised_fn" in names:
        def denoised_fn(v):
            return 0.9 * v + 0.01
        hooks["denoised_fn"] = denoised_fn
    return y_extra, cond_fn, denoised_fn, hooks
