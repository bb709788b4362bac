"""livelyspeaker_b200 - B200-native RAG diffusion sampling path of LivelySpeaker.

Python surface = the reference's (create_model_and_diffusion, load_model_wo_clip,
ClassifierFreeSampleModel, RAG, SpacedDiffusion.p_sample_loop / ddim_sample_loop);
compute = hand-written sm_100a CUDA behind a C ABI (include/livelyspeaker_b200.h).
"""
from ._cabi import MAX_FUSED_STEPS, LsError
from .cfg_sampler import ClassifierFreeSampleModel
from .gaussian_diffusion import GaussianDiffusion, LossType, ModelMeanType, ModelVarType, ReplayNoise, TorchNoise
from .model_util import create_gaussian_diffusion, create_model_and_diffusion, get_model_args, load_model_wo_clip
from .rag import RAG
from .respace import SpacedDiffusion, space_timesteps
from .sag import Decoder_TRANSFORMER
from . import metrics  # noqa: E402,F401

__all__ = ["ClassifierFreeSampleModel", "GaussianDiffusion", "LossType", "ModelMeanType", "ModelVarType",
           "ReplayNoise", "TorchNoise", "create_gaussian_diffusion", "create_model_and_diffusion",
           "get_model_args", "load_model_wo_clip", "RAG", "SpacedDiffusion", "space_timesteps", "Decoder_TRANSFORMER", "metrics"]
