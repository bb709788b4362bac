"""GPU diagnostic: one fused step at B=512 (TED) with LS_FUSED_TIMING=1 -> per-phase cycle stamps of CTA 0."""
import os, sys, types
os.environ["LS_FUSED_TIMING"] = "1"
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
model, diffusion = ls.create_model_and_diffusion(args, "")
model.load_state_dict(sd)
if len(sys.argv) > 2: model.set_impl(sys.argv[2])
cfg = ls.ClassifierFreeSampleModel(model).to(DEV).eval()
eng = model.engine(B)
y = synthetic.synth_cond(dims, B, device=DEV)
eng.set_cond(y, force=True)
g = torch.Generator().manual_seed(4)
x = torch.randn(B, 9, 3, 34, generator=g).to(DEV)
e_c = torch.randn(B, 1, 512, generator=g).to(DEV)
e_u = torch.randn(B, 1, 512, generator=g).to(DEV)
nz = torch.randn(B, 9, 3, 34, generator=g).to(DEV)
p = diffusion.step_params(700, ddim=False, clip_denoised=False)
xp, x0 = torch.empty_like(x), torch.empty_like(x)
for _ in range(2):
    eng.step(p, x, e_c, e_u, nz, y["scale"], xp, x0)
    torch.cuda.synchronize()
