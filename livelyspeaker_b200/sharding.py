"""Batch sharding of the sampling loop over the GPUs of one box.

The path shards by independent units: every clip's trajectory depends only on its own
x_t, conditioning and noise, so ranks take contiguous slices of
the batch, run the loop with NO per-step communication, and one all_gather of the
[B/N, J, D, F] results ends the job.  The reference has no multi-GPU path at all
(scripts/mdm_utils/dist_util.py:18-41 is a no-op).

RNG: `SlicedNoise` makes rank r reproduce rows [lo, hi) of the draws a single-process
run at the global batch would make with the same seed, so a sharded run equals the
1-GPU run bit for bit; the cheaper default (rng='per_rank') is an independent stream per
rank: rank r > 0 re-seeds its generator with seed + r before sampling (fork_seed=False
keeps the caller's own seeding).
"""
import torch
import torch.distributed as dist

_BATCHED_KEYS = ("audio_input", "origin_x", "vid_indices", "scale", "emo", "mask", "lengths", "beat",
                 "text_padded", "inpainting_mask", "inpainted_motion")


def shard_bounds(batch, world_size, rank):
    """Contiguous, balanced split: the first (batch % world) ranks get one extra clip."""
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_cond(y, lo, hi):
    """Slice every per-clip tensor of the cond dict; other entries pass through."""
    out = {}
    for k, v in y.items():
        if k in _BATCHED_KEYS and torch.is_tensor(v) and v.dim() >= 1:
            out[k] = v[lo:hi].clone() if k == "origin_x" else v[lo:hi]
        elif k == "text" and isinstance(v, (list, tuple)):
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


class SlicedNoise:
    """Noise source drawing at the GLOBAL batch and returning rows [lo, hi)."""

    def __init__(self, inner, global_batch, lo, hi):
        self.inner, self.B, self.lo, self.hi = inner, global_batch, lo, hi

    def randn(self, shape, device):
        full = self.inner.randn((self.B,) + tuple(shape[1:]), device)
        return full[self.lo:self.hi]

    def randn_like(self, like):
        # reproduce the layout class of `like` (dense contiguous vs the [F,B,J,D] memory
        # order of a sampler output) at the global batch
        shape = (self.B,) + tuple(like.shape[1:])
        if like.dim() == 4 and not like.is_contiguous():
            full_like = torch.empty(shape[3], shape[0], shape[1], shape[2], device=like.device,
                                    dtype=like.dtype).permute(1, 2, 3, 0)
        else:
            full_like = torch.empty(shape, device=like.device, dtype=like.dtype)
        return self.inner.randn_like(full_like)[self.lo:self.hi]


def all_gather_samples(local, batch, group=None):
    """One all_gather of the per-rank results -> [batch, ...] on every rank.  Ragged
    splits are padded to the largest shard for the collective and trimmed after."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = local.contiguous()      # generic-route results carry the permuted [F,B,J,D] layout: NCCL wants dense tensors
    if local.shape[0] < biggest:
        pad = torch.cat([pad, pad.new_zeros((biggest - pad.shape[0],) + tuple(pad.shape[1:]))], 0)
    bucket = [torch.empty(pad.shape, dtype=pad.dtype, device=pad.device) for _ in range(world)]
    dist.all_gather(bucket, pad, group=group)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bucket, sizes)], 0)


def sample_sharded(sample_fn, model, shape, model_kwargs, diffusion=None, rng="per_rank", group=None,
                   gather=True, **kwargs):
    """Run `sample_fn` (diffusion.p_sample_loop / ddim_sample_loop) on this rank's slice of
    the batch and all_gather the result.  rng='global_slice' reproduces the single-process
    run (needs `diffusion` to swap its noise source for the duration of the call)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = int(shape[0])
    lo, hi = shard_bounds(B, world, rank)
    local_kwargs = dict(model_kwargs)
    local_kwargs["y"] = shard_cond(model_kwargs["y"], lo, hi)
    for k in ("noise", "init_image"):
        if kwargs.get(k) is not None:
            kwargs[k] = kwargs[k][lo:hi]
    saved = saved_model = None
    if rng == "per_rank" and kwargs.pop("fork_seed", True) and rank > 0:
        # The reference seeds every process alike (fixseed, scripts/test_RAG_ted.py:146): identical streams would give
        # every shard the same x_T and noise.  Ranks > 0 therefore advance to their own stream (seed + rank).
        dev = next(model.parameters()).device
        if dev.type == "cuda":
            torch.cuda.manual_seed((torch.cuda.initial_seed() + rank) % (2 ** 63))
        else:
            torch.manual_seed((torch.initial_seed() + rank) % (2 ** 63))
    else:
        kwargs.pop("fork_seed", None)
    if rng == "global_slice":
        if diffusion is None:
            raise ValueError("rng='global_slice' needs the diffusion object")
        saved = diffusion.noise_source
        diffusion.noise_source = SlicedNoise(saved, B, lo, hi)
        # on the generic route the CFG wrapper draws its two style tensors itself (cfg_sampler.py): slice those too
        if hasattr(model, "noise_source"):
            saved_model = (model.noise_source,)
            model.noise_source = SlicedNoise(model.noise_source if model.noise_source is not None else saved, B, lo, hi)
    try:
        local = sample_fn(model, (hi - lo,) + tuple(shape[1:]), model_kwargs=local_kwargs, **kwargs)
    finally:
        if saved is not None:
            diffusion.noise_source = saved
        if saved_model is not None:
            model.noise_source = saved_model[0]
    if not gather:
        return local
    return all_gather_samples(local, B, group)
