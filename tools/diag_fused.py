"""Diagnostic (GPU): fused tcgen05 step vs the SIMT fp32 step, with weight groups rounded
to bf16 to localise which hi/lo term is lost."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic

DEV = "cuda:0"
dims = synthetic.TED
args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1, arch='trans_enc',
                             emb_trans_dec=False, dataset='humanml', lang_model=None, mlpact='silu', diffusion_steps=1000,
                             noise_schedule='cosine', sigma_small=True, lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)


def run(sd, impl, B=4, layers=8):
    a = types.SimpleNamespace(**vars(args)); a.layers = layers
    model, diffusion = ls.create_model_and_diffusion(a, "")
    model.load_state_dict({k: v for k, v in sd.items() if not k.startswith("backbone.mlps.") or int(k.split(".")[2]) < layers})
    model.set_impl(impl)
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
    eng.step(p, x, e_c, e_u, nz, y["scale"], xp, x0)
    torch.cuda.synchronize()
    return x0.clone()


def bf(v):
    return v.bfloat16().float()


base = synthetic.synth_state_dict(dims, seed=1)
variants = {"as is": lambda k: False,
            "W_ch bf16-exact": lambda k: k.endswith("block2.1.weight"),
            "W_tok bf16-exact": lambda k: k.endswith("block1.1.weight"),
            "W_in bf16-exact": lambda k: k == "input_mapping.weight",
            "W_out bf16-exact": lambda k: k == "output_process.poseFinal.weight",
            "all GEMM weights bf16-exact": lambda k: k.endswith(("block2.1.weight", "block1.1.weight")) or k in ("input_mapping.weight", "output_process.poseFinal.weight")}
for layers in (8, 1):
    for name, pick in variants.items():
        sd = {k: (bf(v) if pick(k) else v.clone()) for k, v in base.items()}
        ref = run(sd, "simt", layers=layers)
        for impl in ("tc_bf16x3", "tc_bf16"):
            got = run(sd, impl, layers=layers)
            print("layers=%d  %-30s %-10s max|d| = %.3e   (ref max %.2f)" % (layers, name, impl, float((got - ref).abs().max()), float(ref.abs().max())))
