import sys, types
sys.path.insert(0, "/root/repo")
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1,
                             arch='trans_enc', emb_trans_dec=False, dataset='humanml', lang_model=None,
                             mlpact='silu', diffusion_steps=1000, noise_schedule='cosine', sigma_small=True,
                             lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
model, _ = ls.create_model_and_diffusion(args, "")
ls.load_model_wo_clip(model, synthetic.synth_state_dict(synthetic.TED, seed=1))
model = model.to("cuda:0").eval()
y = synthetic.synth_cond(synthetic.TED, 512, device="cuda:0")
eng = model.engine(512)
for _ in range(3): eng.set_cond(y, force=True)
torch.cuda.synchronize()
