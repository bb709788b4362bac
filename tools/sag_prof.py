import sys, os
sys.path.insert(0, "/root/repo")
import torch, types
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic
DEV="cuda:0"
dec = ls.Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False)
dec.load_state_dict(synthetic.synth_sag_state_dict(seed=3), strict=True)
dec = dec.to(DEV).eval()
B=256
g = torch.Generator().manual_seed(8)
xb, zb = 0.3 * torch.randn(B, 9, 3, 34, generator=g).to(DEV), torch.randn(B, 512, generator=g).to(DEV)
mb = torch.ones(B, 34, dtype=torch.bool, device=DEV)
for _ in range(3): dec({"x": xb, "z": zb, "mask": mb})
torch.cuda.synchronize()
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1,
                             arch='trans_enc', emb_trans_dec=False, dataset='humanml', lang_model=None,
                             mlpact='silu', diffusion_steps=1000, noise_schedule='cosine', sigma_small=True,
                             lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
model, _ = ls.create_model_and_diffusion(args, "")
ls.load_model_wo_clip(model, synthetic.synth_state_dict(synthetic.TED, seed=1))
model = model.to(DEV).eval()
y = synthetic.synth_cond(synthetic.TED, 512, device=DEV)
eng = model.engine(512)
for _ in range(3): eng.set_cond(y, force=True)
torch.cuda.synchronize()
