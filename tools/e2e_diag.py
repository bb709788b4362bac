"""GPU diagnostic: where the end-to-end time of p_sample_loop goes (chunked vs step-by-step)."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic
dev = torch.device("cuda:0")
NAME = sys.argv[1] if len(sys.argv) > 1 else "ted"            # python tools/e2e_diag.py [ted|beat]
dims = synthetic.dims_for(NAME)
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1, arch='trans_enc',
                             emb_trans_dec=False, dataset='humanml', lang_model=None, mlpact='silu', diffusion_steps=1000,
                             noise_schedule='cosine', sigma_small=True, lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
B = 512 if NAME == "ted" else 256
if NAME == "ted":
    model, diffusion = ls.create_model_and_diffusion(args, "")
else:
    from livelyspeaker_b200 import beat_model_util
    args.njoints = 47
    model, diffusion = beat_model_util.create_model_and_diffusion(args, "")
model.load_state_dict(synthetic.synth_state_dict(dims, seed=1))
cfg = ls.ClassifierFreeSampleModel(model).to(dev).eval()
eng = model.engine(B)
y_host = synthetic.synth_cond(dims, B, seed=233)
y_pin = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in y_host.items()}
shape = (B, dims.njoints, dims.nfeats, 34)

def once(n, chunk):
    diffusion.fused_chunk = chunk
    torch.cuda.synchronize(); t0 = time.perf_counter()
    yk = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in y_pin.items()}
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = diffusion.p_sample_loop(cfg, shape, clip_denoised=False, model_kwargs={"y": yk}, skip_timesteps=1000 - n)
    t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    res = out.to("cpu")
    t4 = time.perf_counter()
    print("n=%d chunk=%d: h2d %.1f ms | loop enqueue %.1f ms | loop drain %.1f ms | d2h %.1f ms | total %.1f ms -> %.1f steps/s"
          % (n, chunk, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3, n / (t4 - t0)))

for n, chunk in ((200, 16), (200, 16), (200, 1), (200, 16), (32, 16), (1000, 16)):
    once(n, chunk)
y = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in y_host.items()}
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.set_cond(y, force=True)
    torch.cuda.synchronize(); print("set_cond (WavEncoder + projections) %.2f ms" % ((time.perf_counter() - t0) * 1e3))

if NAME != "ted":
    from livelyspeaker_b200 import gaussian_diffusion as gd
    perm_like = torch.empty(34, B, dims.njoints, dims.nfeats, device=dev).permute(1, 2, 3, 0)
    print("chunk draws served by:", type(gd.chunk_draws(eng, 16, B, 512, perm_like)).__name__)
    sys.exit(0)
# SAG decoder (config 3 front end): time at B=256
dec = ls.Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False)
dec.load_state_dict(synthetic.synth_sag_state_dict(seed=3))
dec = dec.to(dev).eval()
g = torch.Generator().manual_seed(8)
for Bs in (256, 512):
    batch = {"x": (0.3 * torch.randn(Bs, 9, 3, 34, generator=g)).to(dev), "z": torch.randn(Bs, 512, generator=g).to(dev),
             "mask": torch.ones(Bs, 34, dtype=torch.bool, device=dev)}
    dec(dict(batch)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dec(dict(batch))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("SAG decode B=%d: %.2f ms  (%.1f TFLOP/s fp32 algorithmic at 438 MFLOP/clip)" % (Bs, ms, Bs * 438e6 / ms / 1e9))

# which source serves the draws of a full chunk on this box
from livelyspeaker_b200 import gaussian_diffusion as gd
perm_like = torch.empty(34, B, dims.njoints, dims.nfeats, device=dev).permute(1, 2, 3, 0)
srcd = gd.chunk_draws(eng, 16, B, 512, perm_like)
print("chunk draws served by:", type(srcd).__name__, "(fused kernel verified against torch: %s)"
      % eng.graphed_draws(16, B, 512, perm_like, gd._FusedDraws).ok)
