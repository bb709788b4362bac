"""Event-timed FGD feature encoder (ls_pose_features) at B=512: one push_samples encodes two such batches."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from livelyspeaker_b200 import synthetic, embedding_net
DEV = "cuda:0"
net = embedding_net.EmbeddingNet(27, 34).eval()
net.load_state_dict(synthetic.synth_embed_state_dict(seed=5))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
poses = 0.3 * torch.randn(B, 34, 27, device=DEV)
for _ in range(5):
    net(poses)
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        net(poses)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
flops = 2 * B * (32 * 32 * 81 + 30 * 64 * 96 + 14 * 64 * 256 + 12 * 32 * 192 + 384 * 256 + 256 * 128 + 128 * 32 + 2 * 32 * 32)
print("ls_pose_features B=%d: %.1f us per call (host enqueue included), %.2f TFLOP/s fp32" % (B, best * 1e3, flops / best / 1e9))
