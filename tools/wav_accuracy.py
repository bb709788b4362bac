import sys, types
sys.path.insert(0, "/root/repo")
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic, beat_model_util
def args(**kw):
    a = dict(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1, arch='trans_enc', emb_trans_dec=False, dataset='humanml', lang_model=None, mlpact='silu', diffusion_steps=1000, noise_schedule='cosine', sigma_small=True, lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0); a.update(kw); return types.SimpleNamespace(**a)
for name in ("ted", "beat"):
    dims = synthetic.dims_for(name)
    if name == "ted": model, _ = ls.create_model_and_diffusion(args(), "")
    else: model, _ = beat_model_util.create_model_and_diffusion(args(njoints=47), "")
    ls.load_model_wo_clip(model, synthetic.synth_state_dict(dims, seed=1))
    model = model.to("cuda:0").eval()
    for B in (1, 67, 300):
        y = synthetic.synth_cond(dims, B, device="cuda:0")
        model.set_impl("auto"); a = model.engine(B).wav_encoder(y["audio_input"]).clone()
        model.set_impl("simt"); b = model.engine(B).wav_encoder(y["audio_input"]).clone()
        print("%s B=%d: max |tc - fp32| = %.3e, max |fp32| = %.3f, rel-to-tolerance %.3f" % (name, B, float((a-b).abs().max()), float(b.abs().max()), float(((a-b).abs() / (1e-4 + 1e-3*b.abs())).max())))
