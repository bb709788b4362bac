"""GPU diagnostic: per-chunk device time and the gaps between consecutive fused launches of one p_sample_loop
(events recorded around every ls_step_multi call).  usage: python tools/chunk_trace.py [ted|beat]"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic, beat_model_util
dev = torch.device("cuda:0")
NAME = sys.argv[1] if len(sys.argv) > 1 else "ted"
dims = synthetic.dims_for(NAME)
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1, arch='trans_enc',
                             emb_trans_dec=False, dataset='humanml', lang_model=None, mlpact='silu', diffusion_steps=1000,
                             noise_schedule='cosine', sigma_small=True, lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
B = 512 if NAME == "ted" else 256
if NAME == "ted":
    model, diffusion = ls.create_model_and_diffusion(args, "")
else:
    args.njoints = 47
    model, diffusion = beat_model_util.create_model_and_diffusion(args, "")
model.load_state_dict(synthetic.synth_state_dict(dims, seed=1))
cfg = ls.ClassifierFreeSampleModel(model).to(dev).eval()
eng = model.engine(B)
y = synthetic.synth_cond(dims, B, seed=233, device=dev)
shape = (B, dims.njoints, dims.nfeats, 34)
for _ in range(2):
    diffusion.p_sample_loop(cfg, shape, clip_denoised=False, model_kwargs={"y": y}, skip_timesteps=936)
torch.cuda.synchronize()
ev = []
orig = eng.step_multi
def traced(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(*a, **k); e1.record(); ev.append((e0, e1)); return r
eng.step_multi = traced
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
s0.record()
out = diffusion.p_sample_loop(cfg, shape, clip_denoised=False, model_kwargs={"y": y})
s1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
ch = [a.elapsed_time(b) for a, b in ev]
gaps = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1)]
print("%s B=%d: loop %.1f ms on the device (%.1f ms wall), %d launches" % (NAME, B, s0.elapsed_time(s1), (t1 - t0) * 1e3, len(ev)))
print("before the first launch: %.2f ms; first launch %.2f; last (tail) launch %.2f; steady launches median %.3f" %
      (s0.elapsed_time(ev[0][0]), ch[0], ch[-1], sorted(ch[1:-1])[len(ch) // 2]))
print("gaps between launches (draws + host): first %.3f, median %.3f, max %.3f, sum %.2f ms" %
      (gaps[0], sorted(gaps)[len(gaps) // 2], max(gaps), sum(gaps)))
print("sum of launches %.1f ms" % sum(ch))
