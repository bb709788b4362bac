"""Times the SAG decoder (tensor-core and exact-order fp32 path) at B = 256 and the cond precompute at B = 512 / 256."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic

DEV = "cuda:0"


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dec = ls.Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False)
    dec.load_state_dict(synthetic.synth_sag_state_dict(seed=3), strict=True)
    dec = dec.to(DEV).eval()
    for B in (256, 512):
        g = torch.Generator().manual_seed(8)
        xb, zb = 0.3 * torch.randn(B, 9, 3, 34, generator=g).to(DEV), torch.randn(B, 512, generator=g).to(DEV)
        mb = torch.ones(B, 34, dtype=torch.bool, device=DEV)
        res = {}
        for impl in ("tc", "simt"):
            dec.impl = impl
            res[impl] = dec({"x": xb, "z": zb, "mask": mb})["output"].clone()
            print("SAG decode B=%d %s: %.3f ms" % (B, impl, timeit(lambda: dec({"x": xb, "z": zb, "mask": mb}))))
        print("  max |tc - simt| = %.3e (max |simt| %.3f)" % (float((res["tc"] - res["simt"]).abs().max()), float(res["simt"].abs().max())))
    import types
    args = types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1,
                                 arch='trans_enc', emb_trans_dec=False, dataset='humanml', lang_model=None,
                                 mlpact='silu', diffusion_steps=1000, noise_schedule='cosine', sigma_small=True,
                                 lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0)
    model, _ = ls.create_model_and_diffusion(args, "")
    ls.load_model_wo_clip(model, synthetic.synth_state_dict(synthetic.TED, seed=1))
    model = model.to(DEV).eval()
    for B in (256, 512):
        y = synthetic.synth_cond(synthetic.TED, B, device=DEV)
        eng = model.engine(B)
        print("cond precompute B=%d: %.3f ms" % (B, timeit(lambda: eng.set_cond(y, force=True), n=10)))


if __name__ == "__main__":
    main()
