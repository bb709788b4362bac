"""Small-size pass over the round-2 kernels for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic

DEV = "cuda:0"
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1,
                             arch='trans_enc', emb_trans_dec=False, dataset='humanml', lang_model=None,
                             mlpact='silu', diffusion_steps=1000, noise_schedule='cosine', sigma_small=True,
                             lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
model, diffusion = ls.create_model_and_diffusion(args, "ddim100")
ls.load_model_wo_clip(model, synthetic.synth_state_dict(synthetic.TED, seed=1))
cfg = ls.ClassifierFreeSampleModel(model).to(DEV).eval()
B = 3
y = synthetic.synth_cond(synthetic.TED, B, device=DEV)
out = diffusion.ddim_sample_loop(cfg, (B, 9, 3, 34), clip_denoised=False, model_kwargs={"y": y}, skip_timesteps=96)
print("fused loop", tuple(out.shape), bool(torch.isfinite(out).all()))
x = torch.randn(B, 9, 3, 34, device=DEV)
r = diffusion.p_sample_with_grad(cfg, x, torch.tensor([5] * B, device=DEV), clip_denoised=False, model_kwargs={"y": y},
                                 cond_fn=lambda x_, t_, p_, y=None: torch.autograd.grad((p_["pred_xstart"] ** 2).sum(), x_)[0])
print("with_grad step", bool(torch.isfinite(r["sample"]).all()))
model.train()
y["mask"] = torch.ones(B, 34, dtype=torch.bool, device=DEV)
terms, _ = diffusion.training_losses(model, x, torch.tensor([3, 50, 99], device=DEV), model_kwargs={"y": y})
print("training terms", {k: float(v) for k, v in terms.items()})
dec = ls.Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False)
dec.load_state_dict(synthetic.synth_sag_state_dict(seed=3), strict=True)
dec = dec.to(DEV).eval()
o = dec({"x": x, "z": torch.randn(B, 512, device=DEV), "mask": torch.ones(B, 34, dtype=torch.bool, device=DEV)})["output"]
print("sag", bool(torch.isfinite(o).all()))
from livelyspeaker_b200 import embedding_net, metrics
from livelyspeaker_b200.gaussian_diffusion import _FusedDraws
net = embedding_net.EmbeddingNet(27, 34).eval()
net.load_state_dict(synthetic.synth_embed_state_dict(seed=5))
for b in (1, 5):
    f, _, _ = net(0.3 * torch.randn(b, 34, 27, device=DEV))
print("pose features", bool(torch.isfinite(f).all()))
ad, mask = metrics.motion_beats(out)
print("motion beats", int(mask.sum()))
fd = _FusedDraws(16, B, 512, torch.empty(34, B, 9, 3, device=DEV).permute(1, 2, 3, 0))
fd.draw()
print("one-launch draws verified against torch:", fd.ok)
torch.cuda.synchronize()
