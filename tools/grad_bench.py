import sys, types
sys.path.insert(0, "/root/repo")
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic
sys.path.insert(0, "/root/repo/tools")
from sag_bench import timeit
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1,
                             arch='trans_enc', emb_trans_dec=False, dataset='humanml', lang_model=None,
                             mlpact='silu', diffusion_steps=1000, noise_schedule='cosine', sigma_small=True,
                             lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
model, diffusion = ls.create_model_and_diffusion(args, "")
ls.load_model_wo_clip(model, synthetic.synth_state_dict(synthetic.TED, seed=1))
cfg = ls.ClassifierFreeSampleModel(model).to("cuda:0").eval()
for B in (64, 512):
    y = synthetic.synth_cond(synthetic.TED, B, device="cuda:0")
    eng = model.engine(B); eng.set_cond(y, force=True)
    x = torch.randn(B, 9, 3, 34, device="cuda:0"); z = torch.randn(B, 1, 512, device="cuda:0")
    t = torch.full((B,), 500, dtype=torch.long, device="cuda:0")
    g = torch.randn(B, 9, 3, 34, device="cuda:0")
    f = timeit(lambda: eng.cfg_forward_grad(x, t, z, z, y["scale"]), n=5)
    b = timeit(lambda: eng.cfg_backward(g, y["scale"]), n=5)
    tc = timeit(lambda: eng.cfg_forward(x, t, z, z, y["scale"]), n=20)
    cond_fn = lambda x_, t_, p_, y=None: torch.autograd.grad((p_["pred_xstart"] ** 2).sum(), x_)[0]
    s = timeit(lambda: diffusion.p_sample_with_grad(cfg, x, t, clip_denoised=False, model_kwargs={"y": y}, cond_fn=cond_fn), n=5)
    model.train(); y["mask"] = torch.ones(B, 34, dtype=torch.bool, device="cuda:0")
    tl = timeit(lambda: diffusion.training_losses(model, x, t, model_kwargs={"y": y}), n=10)
    model.eval()
    print("B=%d: forward_grad %.2f ms, backward %.2f ms, fused cfg_forward %.2f ms, p_sample_with_grad step %.2f ms, training_losses %.2f ms" % (B, f, b, tc, s, tl))
