"""Multi-rank equality ON HARDWARE: two NCCL ranks, one GPU each, sample_sharded with
rng='global_slice' must reproduce the 1-GPU fused loop bit for bit; the result travels through ONE all_gather.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped elsewhere.  The host-side logic is covered on CPU by test_sharding_gloo.py."""
import os
import socket
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _args():
    return types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=8, cond_mask_prob=0.1,
                                 arch='trans_enc', emb_trans_dec=False, dataset='humanml', lang_model=None, mlpact='silu',
                                 diffusion_steps=1000, noise_schedule='cosine', sigma_small=True, lambda_vel=1.0,
                                 lambda_rcxyz=0.0, lambda_fc=0.0)


def _build(dev):
    import livelyspeaker_b200 as ls
    from livelyspeaker_b200 import synthetic
    dims = synthetic.TED
    model, diffusion = ls.create_model_and_diffusion(_args(), "")
    ls.load_model_wo_clip(model, synthetic.synth_state_dict(dims, seed=1))
    return dims, ls.ClassifierFreeSampleModel(model).to(dev).eval(), diffusion


def _worker(rank, world, port, B, skip, ret):
    import torch.distributed as dist
    from livelyspeaker_b200 import sharding, synthetic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        dims, cfg, diffusion = _build(dev)
        y = synthetic.synth_cond(dims, B, device=dev)
        torch.manual_seed(4321)
        out = sharding.sample_sharded(diffusion.p_sample_loop, cfg, (B, 9, 3, 34), {"y": y}, diffusion=diffusion,
                                      rng="global_slice", clip_denoised=False, skip_timesteps=skip)
        # generic-route result (permuted [F,B,J,D] layout) through the same all_gather: must not trip NCCL's
        # "tensors must be contiguous"
        torch.manual_seed(99)
        out_g = sharding.sample_sharded(diffusion.p_sample_loop, cfg, (B, 9, 3, 34), {"y": y}, diffusion=diffusion,
                                        rng="global_slice", clip_denoised=False, skip_timesteps=997,
                                        denoised_fn=lambda v: v)
        if rank == 0:
            ret["sharded"] = out.cpu()
            ret["generic"] = out_g.cpu()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
@pytest.mark.parametrize("B", [10, 7])
def test_two_nccl_ranks_equal_one_gpu(B):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from livelyspeaker_b200 import synthetic
    skip = 1000 - 35          # 35 steps: two full 16-step launches (one-launch draws) and a 3-step tail
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), B, skip, ret), nprocs=2, join=True)
    dev = torch.device("cuda", 0)
    dims, cfg, diffusion = _build(dev)
    y = synthetic.synth_cond(dims, B, device=dev)
    torch.manual_seed(4321)
    want = diffusion.p_sample_loop(cfg, (B, 9, 3, 34), clip_denoised=False, model_kwargs={"y": y}, skip_timesteps=skip)
    assert torch.equal(ret["sharded"], want.cpu())
    torch.manual_seed(99)
    want_g = diffusion.p_sample_loop(cfg, (B, 9, 3, 34), clip_denoised=False, model_kwargs={"y": y}, skip_timesteps=997,
                                     denoised_fn=lambda v: v)
    assert torch.equal(ret["generic"], want_g.cpu())
