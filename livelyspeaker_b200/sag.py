"""SAG decoder of the LivelySpeaker pipeline.

Mirror of the reference's ``Decoder_TRANSFORMER`` (scripts/model/motionclip_module.py:98-183): same constructor,
same ``state_dict`` keys and shapes (the torch ``nn.TransformerDecoder`` modules are kept as parameter
containers only), same ``forward(batch, use_text_emb=False)`` contract - ``batch['z']`` (CLIP text feature),
``batch['x']`` (pose clip whose first ``n_pre_poses`` frames condition the decoder), ``batch['mask']`` ->
``batch['output']`` / ``batch['txt_output']`` [B, J, D, F] and ``batch['final_z']``.  The math runs through the C
ABI: ``ls_sag_decode_tc`` (csrc/ls_sag_tc.cu: tcgen05 GEMMs, bf16x3, one ``ls_sag`` handle per weight set) by default,
``ls_sag_decode`` (csrc/ls_sag.cu: exact-order fp32, one CTA per clip) with ``decoder.impl = 'simt'``; there is no
PyTorch implementation of it here.
"""
import ctypes
from ctypes import POINTER, c_int32, c_int64, c_void_p

import numpy as np
import torch
import torch.nn as nn

from . import _cabi

SAG_MAX_LAYERS = 8


class LsSagLayer(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("sa_in_wt", "sa_in_b", "sa_out_wt", "sa_out_b", "ca_v_wt", "ca_v_b", "ca_out_wt",
                                        "ca_out_b", "l1_wt", "l1_b", "l2_wt", "l2_b", "n1_w", "n1_b", "n2_w", "n2_b",
                                        "n3_w", "n3_b")]


class LsSagWeights(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("n_layers", "njoints", "nfeats", "n_frames", "n_pre_poses", "latent_dim",
                                       "ff_size", "n_heads")] + \
               [("map_wt", c_void_p), ("map_b", c_void_p), ("pe", c_void_p), ("pe_stride", c_int64),
                ("fin_wt", c_void_p), ("fin_b", c_void_p), ("layer", LsSagLayer * SAG_MAX_LAYERS)]


class PositionalEncoding(nn.Module):
    """Holds the sinusoidal ``pe`` buffer [max_len, 1, d] (motionclip_module.py:12-29); the add happens in the kernel."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0).transpose(0, 1))


class Decoder_TRANSFORMER(nn.Module):
    def __init__(self, modeltype="", njoints=9, nfeats=3, num_frames=34, latent_dim=512, ff_size=1024, num_layers=3,
                 num_heads=4, dropout=0.1, activation="gelu", ablation=None, n_pre_poses=4, use_style=False, **kargs):
        super().__init__()
        self.modeltype, self.njoints, self.nfeats, self.num_frames = modeltype, njoints, nfeats, num_frames
        self.latent_dim, self.ff_size, self.num_layers, self.num_heads = latent_dim, ff_size, num_layers, num_heads
        self.dropout, self.ablation, self.activation = dropout, ablation, activation
        self.input_feats = njoints * nfeats
        self.sequence_pos_encoder = PositionalEncoding(latent_dim, dropout)
        layer = nn.TransformerDecoderLayer(d_model=latent_dim, nhead=num_heads, dim_feedforward=ff_size, dropout=dropout,
                                           activation=activation)
        self.seqTransDecoder = nn.TransformerDecoder(layer, num_layers=num_layers)     # parameter container only
        self.finallayer = nn.Linear(latent_dim, self.input_feats)
        self.mapping = nn.Linear(self.input_feats + 1, latent_dim)
        self.n_pre_poses = n_pre_poses
        self.impl = "tc"              # 'tc': tensor cores (ls_sag_decode_tc); 'simt': exact-order fp32 (ls_sag_decode)
        self._packed = None
        self._tc = None               # (pack key, max_batch, ls_sag*)

    def __del__(self):
        self._drop_tc()

    def _drop_tc(self):
        tc = self.__dict__.get("_tc")
        self.__dict__["_tc"] = None          # not nn.Module.__setattr__: this also runs at interpreter shutdown
        if tc is not None:
            try:
                _cabi.load_library().ls_sag_destroy(tc[2])
            except Exception:      # interpreter shutdown
                pass

    def _tc_handle(self, lib, W, key, bs, dev):
        """The ls_sag handle of the current weights, rebuilt when they changed or the batch outgrew its workspaces."""
        if self._tc is not None and self._tc[0] == key and self._tc[1] >= bs:
            return self._tc[2]
        self._drop_tc()
        lib.ls_sag_create.argtypes = [POINTER(c_void_p), POINTER(LsSagWeights), c_int32, c_int32, c_void_p]
        lib.ls_sag_decode_tc.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        lib.ls_sag_destroy.argtypes = [c_void_p]
        lib.ls_sag_destroy.restype = None
        lib.ls_sag_launch_count.argtypes = [c_void_p]
        lib.ls_sag_launch_count.restype = c_int64
        h = c_void_p()
        rc = lib.ls_sag_create(ctypes.byref(h), ctypes.byref(W), int(bs), dev.index or 0,
                               c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise _cabi.LsError("libls_b200 error %d: %s" % (rc, lib.ls_last_error(None).decode()))
        self._tc = (key, int(bs), h)
        return h

    def launch_count(self):
        """Kernels launched by the tensor-core handle since it was built (0 before the first decode)."""
        return int(_cabi.load_library().ls_sag_launch_count(self._tc[2])) if self._tc is not None else 0

    # ------------------------------------------------------------------ weights -> ls_sag_weights
    def _pack(self, device):
        """Transposed fp32 copies of the matrices + the ctypes struct; rebuilt when a parameter changed."""
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        if activation_of(self) != "gelu":
            raise NotImplementedError("ls_sag_decode implements the GELU feed-forward the reference builds")
        keep = []

        def T(w):            # [out, in] -> [in, out], contiguous, on the device
            t = w.detach().to(device=device, dtype=torch.float32).t().contiguous()
            keep.append(t)
            return t.data_ptr()

        def V(v):
            t = v.detach().to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        d = self.latent_dim
        W = LsSagWeights()
        W.n_layers, W.njoints, W.nfeats, W.n_frames = self.num_layers, self.njoints, self.nfeats, self.num_frames
        W.n_pre_poses, W.latent_dim, W.ff_size, W.n_heads = self.n_pre_poses, d, self.ff_size, self.num_heads
        W.map_wt, W.map_b = T(self.mapping.weight), V(self.mapping.bias)
        W.pe, W.pe_stride = V(self.sequence_pos_encoder.pe[:self.num_frames, 0]), d
        W.fin_wt, W.fin_b = T(self.finallayer.weight), V(self.finallayer.bias)
        if self.num_layers > SAG_MAX_LAYERS:
            raise _cabi.LsError("ls_sag_decode supports at most %d layers" % SAG_MAX_LAYERS)
        for i, lay in enumerate(self.seqTransDecoder.layers):
            L = W.layer[i]
            L.sa_in_wt, L.sa_in_b = T(lay.self_attn.in_proj_weight), V(lay.self_attn.in_proj_bias)
            L.sa_out_wt, L.sa_out_b = T(lay.self_attn.out_proj.weight), V(lay.self_attn.out_proj.bias)
            L.ca_v_wt = T(lay.multihead_attn.in_proj_weight[2 * d:])
            L.ca_v_b = V(lay.multihead_attn.in_proj_bias[2 * d:])
            L.ca_out_wt, L.ca_out_b = T(lay.multihead_attn.out_proj.weight), V(lay.multihead_attn.out_proj.bias)
            L.l1_wt, L.l1_b = T(lay.linear1.weight), V(lay.linear1.bias)
            L.l2_wt, L.l2_b = T(lay.linear2.weight), V(lay.linear2.bias)
            L.n1_w, L.n1_b = V(lay.norm1.weight), V(lay.norm1.bias)
            L.n2_w, L.n2_b = V(lay.norm2.weight), V(lay.norm2.bias)
            L.n3_w, L.n3_b = V(lay.norm3.weight), V(lay.norm3.bias)
        self._packed = (key, (W, keep))
        return self._packed[1]

    def forward(self, batch, use_text_emb=False):
        if self.training:
            raise NotImplementedError("the SAG decoder is an inference path here: call .eval() first")
        z, mask = batch["z"], batch["mask"]
        if use_text_emb:
            z = batch["clip_text_emb"]
        bs, nframes = mask.shape
        if nframes != self.num_frames:
            raise _cabi.LsError("mask has %d frames, the decoder was built for %d" % (nframes, self.num_frames))
        batch['final_z'] = z.clone()
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _cabi.LsError("the SAG decoder runs on a CUDA sm_100 device only (no CPU path): move it with .to('cuda')")
        lib = _cabi.load_library()
        lib.ls_sag_decode.argtypes = [POINTER(LsSagWeights), c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        W, _keep = self._pack(dev)
        x = batch["x"].to(dev).float().reshape(bs, self.input_feats, nframes).contiguous()
        zf = z.to(dev).float().contiguous()
        m8 = mask.to(dev).to(torch.uint8).contiguous()
        out = torch.empty(bs, self.njoints, self.nfeats, nframes, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            args = (bs, c_void_p(x.data_ptr()), c_void_p(zf.data_ptr()), c_void_p(m8.data_ptr()),
                    c_void_p(out.data_ptr()), c_void_p(torch.cuda.current_stream().cuda_stream))
            if self.impl == "simt":
                rc = lib.ls_sag_decode(ctypes.byref(W), *args)
            elif self.impl == "tc":
                rc = lib.ls_sag_decode_tc(self._tc_handle(lib, W, self._packed[0], bs, dev), *args)
            else:
                raise ValueError("impl must be 'tc' or 'simt'")
        if rc != 0:
            raise _cabi.LsError("libls_b200 error %d: %s" % (rc, lib.ls_last_error(None).decode()))
        # the reference returns output.permute(1, 2, 3, 0) of a [F,B,J,D] tensor (motionclip_module.py:181): same
        # values, memory order [F,B,J,D] - it decides the element order of a later randn_like(init_image)
        batch["txt_output" if use_text_emb else "output"] = _cabi._ref_layout(out)
        return batch


def activation_of(dec):
    return dec.activation if isinstance(dec.activation, str) else getattr(dec.activation, "__name__", "gelu")
