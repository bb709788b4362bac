import os, sys, types
sys.path.insert(0, "/root/repo")
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic, gaussian_diffusion as gd
dev = torch.device("cuda:0")
dims = synthetic.TED
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1, arch='trans_enc',
                             emb_trans_dec=False, dataset='humanml', lang_model=None, mlpact='silu', diffusion_steps=1000,
                             noise_schedule='cosine', sigma_small=True, lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
B = 512
model, diffusion = ls.create_model_and_diffusion(args, "")
model.load_state_dict(synthetic.synth_state_dict(dims, seed=1))
cfg = ls.ClassifierFreeSampleModel(model).to(dev).eval()
eng = model.engine(B)
perm_like = torch.empty(34, B, 9, 3, device=dev).permute(1, 2, 3, 0)
f = gd._FusedDraws(16, B, 512, perm_like)
print("fused ok:", f.ok)
# detailed comparison
order = [t for k in range(16) for t in (f.eps_c[k], f.eps_u[k], f.nz[k])]
state = torch.cuda.get_rng_state(dev)
want = [torch.randn_like(t) for t in order]
end_t = torch.cuda.get_rng_state(dev)
torch.cuda.set_rng_state(state, dev)
f._launch()
end_o = torch.cuda.get_rng_state(dev)
for i, (a, b) in enumerate(zip(order, want)):
    if not torch.equal(a, b):
        d = (a - b).abs()
        print("tensor", i, "numel", a.numel(), "mismatch: max", float(d.max()), "frac", float((d > 0).float().mean()), a.flatten()[:4].tolist(), b.flatten()[:4].tolist())
        break
else:
    print("all tensors equal")
print("end state equal:", torch.equal(end_t, end_o), end_t[-16:].tolist(), end_o[-16:].tolist())
g = torch.cuda.default_generators[0]
print("seed", g.initial_seed(), "offset", g.get_offset())
import time
for name, src in (("fused", f), ("graph", gd._GraphedDraws(16, B, 512, perm_like))):
    src.draw(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): src.draw()
    e1.record(); torch.cuda.synchronize()
    print(name, "draw of one chunk: %.3f ms" % (e0.elapsed_time(e1) / 20))
