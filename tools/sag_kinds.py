"""Real (back-to-back, CUDA-event) time of each kernel kind of the tensor-core SAG decode: run with LS_SAG_MASK=<bit>."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KINDS = ["queries", "cross", "gemm QKV x3", "attention x3", "gemm out+LN1+LN2 x3", "gemm FFN1+GELU x3", "gemm FFN2+LN3 x3", "final"]
if len(sys.argv) > 1:
    import torch
    import livelyspeaker_b200 as ls
    from livelyspeaker_b200 import synthetic
    from sag_bench import timeit
    dec = ls.Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False)
    dec.load_state_dict(synthetic.synth_sag_state_dict(seed=3), strict=True)
    dec = dec.to("cuda:0").eval()
    B = int(sys.argv[1])
    g = torch.Generator().manual_seed(8)
    batch = {"x": 0.3 * torch.randn(B, 9, 3, 34, generator=g).cuda(), "z": torch.randn(B, 512, generator=g).cuda(),
             "mask": torch.ones(B, 34, dtype=torch.bool, device="cuda:0")}
    print("%.1f" % (1000 * timeit(lambda: dec(batch), n=50)))
else:
    for B in (256,):
        for i, k in enumerate(KINDS + ["all"]):
            env = dict(os.environ, LS_SAG_MASK=str(1 << i if i < 8 else 255))
            out = subprocess.run([sys.executable, __file__, str(B)], env=env, capture_output=True, text=True, cwd=os.path.join(ROOT, "tools"))
            print("B=%d %-24s %s us" % (B, k, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]))
