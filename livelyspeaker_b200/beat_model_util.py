"""BEAT twin of model_util (scripts_beat/mdm_utils/model_util.py): njoints from args, nfeats 6,
36 tokens (style + emotion), and the BEAT tree's sampler quirks."""
from .model_util import load_model_wo_clip  # noqa: F401
from . import model_util as _mu


def create_model_and_diffusion(args, timestep_respacing=''):
    return _mu.create_model_and_diffusion(args, timestep_respacing, variant='beat')


def get_model_args(args):
    return _mu.get_model_args(args, 'beat')


def create_gaussian_diffusion(args, timestep_respacing=''):
    return _mu.create_gaussian_diffusion(args, timestep_respacing, 'beat')
