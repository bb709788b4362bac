"""Pose autoencoder of the TED evaluation - the encoder side, for the FGD / diversity features.

Mirror of the reference's ``EmbeddingNet`` (scripts/model/embedding_net.py:266-280) as the evaluator uses it
(scripts/model/ted_evaluator.py:16-24, 35-41): same constructor, the same ``state_dict`` key set and shapes - so the
``gen_dict`` of ``gesture_autoencoder_checkpoint_best.bin`` loads strictly - and the same
``forward(poses, variational_encoding=False) -> (feat, mu, logvar)`` contract.  The torch layers are parameter
containers only: the encoder's arithmetic (four convolutions, BatchNorm on running statistics, three linear layers,
the two heads) runs in one CUDA kernel behind the C ABI (``ls_pose_features``, csrc/ls_pose_feat.cu); there is no
PyTorch or CPU implementation of it here.  The decoder (``PoseDecoderConv``, :166-216) is kept for its parameters -
the evaluator never runs it (its reconstruction-error lines are commented out, ted_evaluator.py:43-46).
"""
import ctypes
from ctypes import c_float, c_int32, c_void_p

import torch
import torch.nn as nn

from . import _cabi

_PTRS = ("c1_w", "c1_b", "c1_scale", "c1_shift", "c2_w", "c2_b", "c2_scale", "c2_shift", "c3_w", "c3_b", "c3_scale",
         "c3_shift", "c4_w", "c4_b", "f1_wt", "f1_b", "f1_scale", "f1_shift", "f2_wt", "f2_b", "f2_scale", "f2_shift",
         "f3_wt", "f3_b", "mu_wt", "mu_b", "lv_wt", "lv_b")


class LsPoseEncoderWeights(ctypes.Structure):
    """``ls_pose_encoder_weights`` of include/livelyspeaker_b200.h."""
    _fields_ = [("pose_dim", c_int32), ("n_frames", c_int32), ("slope_conv", c_float), ("slope_fc", c_float)] + \
               [(n, c_void_p) for n in _PTRS]


def _conv_bn_act(c_in, c_out, downsample=False):
    """embedding_net.py:15-37 with batchnorm=True: kernel 3 / stride 1, or kernel 4 / stride 2 when downsampling."""
    k, s = (4, 2) if downsample else (3, 1)
    return nn.Sequential(nn.Conv1d(c_in, c_out, kernel_size=k, stride=s), nn.BatchNorm1d(c_out), nn.LeakyReLU(0.2, True))


class PoseEncoderConv(nn.Module):
    def __init__(self, length, dim):
        super().__init__()
        if length != 34:
            raise NotImplementedError("the evaluator's encoder is built for 34-frame clips (384-wide first linear layer)")
        self.length, self.dim = length, dim
        self.net = nn.Sequential(_conv_bn_act(dim, 32), _conv_bn_act(32, 64), _conv_bn_act(64, 64, True), nn.Conv1d(64, 32, 3))
        # LeakyReLU(True): the reference passes True as the negative slope, which makes these two the identity
        self.out_net = nn.Sequential(nn.Linear(384, 256), nn.BatchNorm1d(256), nn.LeakyReLU(True), nn.Linear(256, 128),
                                     nn.BatchNorm1d(128), nn.LeakyReLU(True), nn.Linear(128, 32))
        self.fc_mu = nn.Linear(32, 32)
        self.fc_logvar = nn.Linear(32, 32)
        self._packed = None

    def _pack(self, device):
        """Device copies in the layout of ``ls_pose_encoder_weights`` (BatchNorm folded in fp64); rebuilt when a
        parameter or running statistic changed."""
        tensors = list(self.parameters()) + [b for b in self.buffers()]
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in tensors)
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        keep = {}

        def put(name, t):
            keep[name] = t.detach().to(device=device, dtype=torch.float32).contiguous()

        def fold(prefix, bn):
            s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            put(prefix + "_scale", s)
            put(prefix + "_shift", bn.bias.detach().double() - bn.running_mean.detach().double() * s)

        for i in range(3):
            put("c%d_w" % (i + 1), self.net[i][0].weight)
            put("c%d_b" % (i + 1), self.net[i][0].bias)
            fold("c%d" % (i + 1), self.net[i][1])
        put("c4_w", self.net[3].weight)
        put("c4_b", self.net[3].bias)
        for name, lin, bn in (("f1", 0, 1), ("f2", 3, 4)):
            put(name + "_wt", self.out_net[lin].weight.detach().t())
            put(name + "_b", self.out_net[lin].bias)
            fold(name, self.out_net[bn])
        put("f3_wt", self.out_net[6].weight.detach().t())
        put("f3_b", self.out_net[6].bias)
        put("mu_wt", self.fc_mu.weight.detach().t())
        put("mu_b", self.fc_mu.bias)
        put("lv_wt", self.fc_logvar.weight.detach().t())
        put("lv_b", self.fc_logvar.bias)
        W = LsPoseEncoderWeights()
        W.pose_dim, W.n_frames = self.dim, self.length
        W.slope_conv = float(self.net[0][2].negative_slope)
        W.slope_fc = float(self.out_net[2].negative_slope)
        for n in _PTRS:
            setattr(W, n, keep[n].data_ptr())
        self._packed = (key, (W, keep))
        return self._packed[1]

    def forward(self, poses, variational_encoding):
        """poses [B, length, dim] -> (z, mu, logvar), each [B, 32] (embedding_net.py:64-79)."""
        if self.training:
            raise NotImplementedError("BatchNorm on batch statistics (training mode) is not built; call .eval() like "
                                      "the evaluator does (ted_evaluator.py:21-22)")
        if poses.device.type != "cuda":
            raise _cabi.LsError("PoseEncoderConv runs on a CUDA sm_100 device only (no CPU path)")
        if poses.dim() != 3 or poses.shape[1] != self.length or poses.shape[2] != self.dim:
            raise ValueError("poses must be [B, %d, %d], got %s" % (self.length, self.dim, tuple(poses.shape)))
        lib = _cabi.load_library()
        W, _keep = self._pack(poses.device)
        x = poses.detach().float().contiguous()
        B = x.shape[0]
        mu = torch.empty(B, 32, dtype=torch.float32, device=x.device)
        logvar = torch.empty(B, 32, dtype=torch.float32, device=x.device)
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(x.device):
            rc = lib.ls_pose_features(ctypes.byref(W), B, c_void_p(x.data_ptr()), c_void_p(mu.data_ptr()),
                                      c_void_p(logvar.data_ptr()), dev, c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise _cabi.LsError("libls_b200 error %d: %s" % (rc, lib.ls_last_error(None).decode()))
        if variational_encoding:                # reparameterize (:9-12): one randn_like draw from the global generator
            std = torch.exp(0.5 * logvar)
            z = mu + torch.randn_like(std) * std
        else:
            z = mu
        return z, mu, logvar


class PoseDecoderConv(nn.Module):
    """Parameter container of the autoencoder's decoder (embedding_net.py:166-216); the evaluator never runs it."""

    def __init__(self, length, dim, use_pre_poses=False):
        super().__init__()
        if length != 34 or use_pre_poses:
            raise NotImplementedError("only the 34-frame decoder without pre-poses that EmbeddingNet builds")
        self.pre_net = nn.Sequential(nn.Linear(32, 64), nn.BatchNorm1d(64), nn.LeakyReLU(True), nn.Linear(64, 136))
        self.net = nn.Sequential(nn.ConvTranspose1d(4, 32, 3), nn.BatchNorm1d(32), nn.LeakyReLU(0.2, True),
                                 nn.ConvTranspose1d(32, 32, 3), nn.BatchNorm1d(32), nn.LeakyReLU(0.2, True),
                                 nn.Conv1d(32, 32, 3), nn.Conv1d(32, dim, 3))

    def forward(self, feat, pre_poses=None):
        raise NotImplementedError("pose reconstruction is off the evaluation path (ted_evaluator.py:43-46) and not built")


class EmbeddingNet(nn.Module):
    def __init__(self, pose_dim, n_frames):
        super().__init__()
        self.pose_encoder = PoseEncoderConv(n_frames, pose_dim)
        self.decoder = PoseDecoderConv(n_frames, pose_dim)

    def forward(self, poses, variational_encoding=False):
        return self.pose_encoder(poses, variational_encoding)

    def freeze_pose_nets(self):
        for p in self.parameters():
            p.requires_grad = False
