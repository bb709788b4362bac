"""Factory + checkpoint loader with the reference's signatures
(scripts/mdm_utils/model_util.py:5-74 for TED; scripts_beat/mdm_utils/model_util.py
for BEAT, selected with ``variant='beat'`` or by importing `beat_model_util`).
"""
from . import gaussian_diffusion as gd
from .rag import RAG
from .respace import SpacedDiffusion, space_timesteps


def load_model_wo_clip(model, state_dict):
    """strict=False load; no unexpected keys; only `clip_model.*` may be missing."""
    missing_keys, unexpected_keys = model.load_state_dict(state_dict, strict=False)
    print("missing_keys", missing_keys)
    print("unexpected_keys", unexpected_keys)
    assert len(unexpected_keys) == 0
    assert all([k.startswith('clip_model.') for k in missing_keys])


def create_model_and_diffusion(args, timestep_respacing='', variant='ted'):
    model = RAG(**get_model_args(args, variant))
    diffusion = create_gaussian_diffusion(args, timestep_respacing, variant)
    return model, diffusion


def get_model_args(args, variant='ted'):
    beat = variant == 'beat'
    return {'modeltype': '', 'njoints': args.njoints if beat else 9, 'nfeats': 6 if beat else 3,
            'num_actions': 1370, 'translation': True, 'pose_rep': 'rot6d', 'glob': True, 'glob_rot': True,
            'latent_dim': args.latent_dim, 'ff_size': 1024 if beat else args.ff_size, 'num_layers': args.layers,
            'num_heads': 4, 'dropout': 0.1, 'activation': "gelu", 'data_rep': 'vec_dir',
            'cond_mode': args.mdm_condm, 'cond_mask_prob': args.cond_mask_prob, 'action_emb': 'tensor',
            'arch': args.arch, 'emb_trans_dec': args.emb_trans_dec, 'clip_version': 'ViT-B/32',
            'dataset': args.dataset, 'lang_model': args.lang_model, 'mlpact': 'silu' if beat else args.mlpact,
            'n_pre_emb': 2 if beat else 1}


def create_gaussian_diffusion(args, timestep_respacing='', variant='ted'):
    steps = args.diffusion_steps
    betas = gd.get_named_beta_schedule(args.noise_schedule, steps, 1.)
    if not timestep_respacing:
        timestep_respacing = [steps]
    diffusion = SpacedDiffusion(
        use_timesteps=sorted(space_timesteps(steps, timestep_respacing)),
        betas=betas,
        model_mean_type=gd.ModelMeanType.START_X,        # "we always predict x_start"
        model_var_type=gd.ModelVarType.FIXED_SMALL if args.sigma_small else gd.ModelVarType.FIXED_LARGE,
        loss_type=gd.LossType.HUBER,
        rescale_timesteps=False,
        lambda_vel=args.lambda_vel, lambda_rcxyz=args.lambda_rcxyz, lambda_fc=args.lambda_fc)
    if variant == 'beat':
        # the BEAT tree's sampler differs in four observable ways
        # (scripts_beat/diffusion/gaussian_diffusion.py:319, 665, 700-704, 913-914)
        diffusion.dump_key = "sample"
        diffusion.allow_ddim_const_noise = False
        diffusion.inpaint_noised = False
        diffusion.const_noise_init = False   # scripts_beat/...:700-704: x_T is NOT repeated under const_noise
        diffusion.training_returns_pred = False   # scripts_beat/...:1388 returns terms only (TED: (terms, pred), :1396-1401)
    return diffusion
