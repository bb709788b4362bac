"""GPU time of the whole tensor-core SAG decode without host overhead: the decode captured in a CUDA graph and replayed."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import livelyspeaker_b200 as ls
from livelyspeaker_b200 import synthetic

dec = ls.Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False)
dec.load_state_dict(synthetic.synth_sag_state_dict(seed=3), strict=True)
dec = dec.to("cuda:0").eval()
for B in (256, 512):
    g = torch.Generator().manual_seed(8)
    batch = {"x": 0.3 * torch.randn(B, 9, 3, 34, generator=g).cuda(), "z": torch.randn(B, 512, generator=g).cuda(),
             "mask": torch.ones(B, 34, dtype=torch.bool, device="cuda:0")}
    for _ in range(3):
        dec(dict(batch))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        dec(dict(batch))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("B=%d host time per decode call (no sync): %.1f us" % (B, (t1 - t0) / 20 * 1e6))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        dec(dict(batch))
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = dec(dict(batch))["output"]
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        gr.replay()
    a.record()
    for _ in range(50):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    print("B=%d graph replay: %.1f us per decode" % (B, a.elapsed_time(b) / 50 * 1000))
