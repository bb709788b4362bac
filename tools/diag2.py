import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic
DEV = "cuda:0"
dims = synthetic.TED
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1, arch='trans_enc',
                             emb_trans_dec=False, dataset='humanml', lang_model=None, mlpact='silu', diffusion_steps=1000,
                             noise_schedule='cosine', sigma_small=True, lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
sd = synthetic.synth_state_dict(dims, seed=1)
def run(impl, B, t, reps=1):
    model, diffusion = ls.create_model_and_diffusion(args, "")
    model.load_state_dict(sd)
    model.set_impl(impl)
    cfg = ls.ClassifierFreeSampleModel(model).to(DEV).eval()
    eng = model.engine(B)
    y = synthetic.synth_cond(dims, B, device=DEV)
    eng.set_cond(y, force=True)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, 9, 3, 34, generator=g).to(DEV)
    e_c = torch.randn(B, 1, 512, generator=g).to(DEV)
    e_u = torch.randn(B, 1, 512, generator=g).to(DEV)
    nz = torch.randn(B, 9, 3, 34, generator=g).to(DEV)
    p = diffusion.step_params(t, ddim=False, clip_denoised=False)
    outs = []
    for r in range(reps):
        xp, x0 = torch.empty_like(x), torch.empty_like(x)
        eng.step(p, x, e_c, e_u, nz, y["scale"], xp, x0)
        torch.cuda.synchronize()
        outs.append(x0.clone())
    return outs
for mask in (0, 64, 64+128, 63+128):
  os.environ["LS_DBG_MASK"] = str(mask)
  for B, t in ((2, 700),):
    ref = run("simt", B, t)[0]
    for impl in ("tc_bf16x3", "tc_bf16"):
        outs = run(impl, B, t, reps=3)
        per_clip = [(o - ref).abs().flatten(1).max(1).values.tolist() for o in outs]
        print("mask", mask, B, t, impl, ["%.1e" % v for v in per_clip[0]], "rep1:", ["%.1e" % v for v in per_clip[1]], "rep2==rep1:", bool(torch.equal(outs[1], outs[2])))
