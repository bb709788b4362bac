"""Event-timed cond precompute (WavEncoder + conditioning projections) at B=512, TED shape."""
import sys, types
sys.path.insert(0, "/root/repo")
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic
DEV = "cuda:0"
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1,
                             arch='trans_enc', emb_trans_dec=False, dataset='humanml', lang_model=None,
                             mlpact='silu', diffusion_steps=1000, noise_schedule='cosine', sigma_small=True,
                             lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
model, _ = ls.create_model_and_diffusion(args, "")
ls.load_model_wo_clip(model, synthetic.synth_state_dict(synthetic.TED, seed=1))
model = model.to(DEV).eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
y = synthetic.synth_cond(synthetic.TED, B, device=DEV)
eng = model.engine(B)
for _ in range(5):
    eng.set_cond(y, force=True)
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.set_cond(y, force=True)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print("cond precompute B=%d: %.3f ms per call (best of 5 x 20)" % (B, best))
