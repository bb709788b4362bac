"""ctypes binding of libls_b200.so (include/livelyspeaker_b200.h) and the `Engine`
object the Python mirror of the reference interface drives.

There is deliberately NO fallback: if the shared library is missing or the
device is not a CUDA sm_100 GPU, every compute entry point raises.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# LS_B200_LIB selects a diagnostic build of the same library (tools/build_variants.sh); never a fallback.
LIB_PATH = os.environ.get("LS_B200_LIB") or os.path.join(_HERE, "csrc", "libls_b200.so")

LS_IMPL_AUTO, LS_IMPL_SIMT, LS_IMPL_TC_BF16X3, LS_IMPL_TC_BF16 = 0, 1, 2, 3
IMPL_NAMES = {"auto": 0, "simt": 1, "tc_bf16x3": 2, "tc_bf16": 3}

EXPORTS = ["ls_abi_version", "ls_last_error", "ls_create", "ls_destroy", "ls_load_weight",
           "ls_finalize_weights", "ls_set_impl", "ls_get_impl", "ls_precompute_cond", "ls_wav_encoder",
           "ls_model_forward", "ls_model_forward_train", "ls_huber_terms", "ls_cfg_forward", "ls_cfg_forward_grad", "ls_cfg_backward", "ls_step", "ls_step_multi", "ls_sag_decode", "ls_sag_create", "ls_sag_decode_tc",
           "ls_sag_launch_count", "ls_sag_destroy", "ls_randn_torch_compat", "ls_q_sample", "ls_launch_count",
           "ls_debug_buffer", "ls_debug_hidden", "ls_motion_beats", "ls_beat_align", "ls_pose_features", "ls_vb_terms"]


class LsConfig(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("njoints", "nfeats", "n_frames", "n_pre_emb", "latent_dim", "n_layers",
                                       "audio_len", "n_speakers", "n_emotions", "max_batch", "max_timestep",
                                       "device")]


class LsStepParams(ctypes.Structure):
    _fields_ = [("mode", c_int32), ("t_model", c_int32), ("clip_denoised", c_int32), ("add_noise", c_int32),
                ("c", c_float * 8)]


class LsStepIO(ctypes.Structure):
    _fields_ = [("eps_cond", c_void_p), ("eps_uncond", c_void_p), ("noise", c_void_p),
                ("noise_sb", c_int64), ("noise_sj", c_int64), ("noise_sf", c_int64),
                ("x_prev", c_void_p), ("pred_x0", c_void_p)]


MAX_FUSED_STEPS = 16     # LS_MAX_FUSED_STEPS


class LsError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen libls_b200.so once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LsError("%s not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.ls_abi_version.restype = c_int32
    lib.ls_last_error.restype = c_char_p
    lib.ls_last_error.argtypes = [c_void_p]
    lib.ls_create.argtypes = [POINTER(c_void_p), POINTER(LsConfig)]
    lib.ls_destroy.argtypes = [c_void_p]
    lib.ls_destroy.restype = None
    lib.ls_load_weight.argtypes = [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32, c_void_p]
    lib.ls_finalize_weights.argtypes = [c_void_p, c_void_p]
    lib.ls_set_impl.argtypes = [c_void_p, c_int32]
    lib.ls_get_impl.argtypes = [c_void_p]
    lib.ls_precompute_cond.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int32, c_void_p]
    lib.ls_wav_encoder.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]
    lib.ls_model_forward.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]
    lib.ls_cfg_forward.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p]
    lib.ls_step.argtypes = [c_void_p, c_int32, POINTER(LsStepParams), c_void_p, c_void_p, c_void_p, c_void_p,
                            c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ls_step_multi.argtypes = [c_void_p, c_int32, c_int32, POINTER(LsStepParams), POINTER(LsStepIO), c_void_p,
                                  c_void_p, c_void_p]
    lib.ls_randn_torch_compat.argtypes = [c_int32, POINTER(c_void_p), POINTER(c_int64), ctypes.c_uint64, ctypes.c_uint64,
                                          POINTER(ctypes.c_uint64), c_int32, c_void_p]
    lib.ls_q_sample.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p]
    lib.ls_launch_count.argtypes = [c_void_p]
    lib.ls_launch_count.restype = c_int64
    lib.ls_debug_buffer.argtypes = [c_void_p, c_int32, c_void_p, c_int64, POINTER(c_int64), c_void_p]
    lib.ls_debug_hidden.argtypes = [c_void_p, c_int32, c_void_p]
    lib.ls_motion_beats.argtypes = [c_int32, c_int32, c_int32, c_void_p, POINTER(c_float), POINTER(c_int32),
                                    POINTER(c_float), c_int32, c_float, c_void_p, c_void_p, c_int32, c_void_p]
    lib.ls_beat_align.argtypes = [c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_float, c_float, c_void_p,
                                  c_void_p, c_void_p, c_int32, c_void_p]
    if lib.ls_abi_version() != 1:
        raise LsError("ABI version mismatch: library %d, binding 1" % lib.ls_abi_version())
    _lib = lib
    return lib


def _f32(t, device):
    """Dense fp32 tensor on `device` (pinned host tensors are copied asynchronously)."""
    if t.device != device:
        t = t.to(device, non_blocking=True)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ref_layout(out):
    """Give a dense [B,J,D,F] result the strides the reference's OutputProcess produces
    (memory order [F,B,J,D], RAG.py:205-211): torch.randn_like on tensors derived from it
    then draws in the same order as in the reference (DESIGN.md "RNG layout")."""
    return out.permute(3, 0, 1, 2).contiguous().permute(1, 2, 3, 0)


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One ls_handle: weights + workspaces for one model on one GPU."""

    def __init__(self, dims, device, max_batch=512, max_timestep=1000, n_speakers=1400, n_emotions=0):
        device = torch.device(device)
        if device.type != "cuda":
            raise LsError("livelyspeaker_b200 computes on CUDA sm_100a only (got device %s); "
                          "there is no CPU path" % device)
        self.lib = load_library()
        self.dims = dims
        self.device = device
        self.max_batch = int(max_batch)
        self.max_timestep = int(max_timestep)
        self.cfg = LsConfig(dims.njoints, dims.nfeats, 34, dims.n_pre_emb, dims.latent_dim, dims.layers,
                            dims.audio_len, n_speakers, n_emotions, self.max_batch, self.max_timestep,
                            device.index if device.index is not None else torch.cuda.current_device())
        self.h = c_void_p()
        rc = self.lib.ls_create(byref(self.h), byref(self.cfg))
        if rc != 0:
            raise LsError("ls_create failed (%d): %s" % (rc, self.lib.ls_last_error(None).decode()))
        self._weights_sig = None
        self._cond_sig = None
        self._cond_keep = None
        self._impl = LS_IMPL_AUTO

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                self.lib.ls_destroy(self.h)
                self.h = c_void_p()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise LsError("libls_b200 error %d: %s" % (rc, self.lib.ls_last_error(self.h).decode()))

    # ---- weights -----------------------------------------------------------------
    def load_state_dict(self, sd):
        """sd: reference-named tensors (any device); uploads + derives kernel layouts."""
        with torch.cuda.device(self.device):
            keep = []
            for key, val in sd.items():
                if not torch.is_tensor(val):
                    continue
                if key.endswith("sequence_pos_encoder.pe") and key != "backbone.embed_timestep.sequence_pos_encoder.pe":
                    continue
                t = _f32(val.detach(), self.device)
                keep.append(t)
                shape = (c_int64 * t.dim())(*t.shape)
                self._check(self.lib.ls_load_weight(self.h, key.encode(), c_void_p(t.data_ptr()), shape, t.dim(),
                                                    _stream()))
            self._check(self.lib.ls_finalize_weights(self.h, _stream()))
            self._check(self.lib.ls_set_impl(self.h, self._impl))
            torch.cuda.current_stream().synchronize()   # `keep` may be freed after this
        self._cond_sig = None

    def set_impl(self, impl):
        impl = IMPL_NAMES[impl] if isinstance(impl, str) else int(impl)
        self._check(self.lib.ls_set_impl(self.h, impl))
        self._impl = impl

    def get_impl(self):
        v = self.lib.ls_get_impl(self.h)
        return {v_: k for k, v_ in IMPL_NAMES.items()}[v]

    def launch_count(self):
        return int(self.lib.ls_launch_count(self.h))

    # ---- conditioning ---------------------------------------------------------------
    @staticmethod
    def _sig(y):
        sig = []
        for k in ("audio_input", "origin_x", "vid_indices", "emo"):
            t = y.get(k)
            sig.append(None if t is None else (t.data_ptr(), t._version, tuple(t.shape), str(t.device)))
        return tuple(sig)

    def set_cond(self, y, force=False):
        """ls_precompute_cond on y (cached on tensor identity + version).  Zeroes
        y['origin_x'][..., 4:] in the caller's tensor like RAG.py:110."""
        sig = self._sig(y)
        if not force and sig == self._cond_sig:
            return
        audio = _f32(y["audio_input"], self.device)
        B = audio.shape[0]
        if B > self.max_batch:
            raise LsError("batch %d exceeds the engine's max_batch %d" % (B, self.max_batch))
        d = self.dims
        # The kernels stride the audio by cfg.audio_len and origin_x by J*D*34: any other shape would silently read
        # the wrong rows.  (The reference WavEncoder accepts any length, but RAG.forward's torch.cat only works when
        # it yields exactly 34 frames - audio_enc.py:22-25, RAG.py:112.)
        if audio.dim() != 2 or audio.shape[1] != d.audio_len:
            raise LsError("y['audio_input'] must be [B,%d] for this model (got %s)" % (d.audio_len, tuple(audio.shape)))
        ox_user = y["origin_x"]
        if tuple(ox_user.shape) != (B, d.njoints, d.nfeats, 34):
            raise LsError("y['origin_x'] must be [%d,%d,%d,34] (got %s)" % (B, d.njoints, d.nfeats, tuple(ox_user.shape)))
        if y["vid_indices"].numel() != B:
            raise LsError("y['vid_indices'] must hold one speaker id per clip (%d), got %d" % (B, y["vid_indices"].numel()))
        inplace = (ox_user.device == self.device and ox_user.dtype == torch.float32 and ox_user.is_contiguous())
        ox = ox_user if inplace else _f32(ox_user, self.device).clone()
        vid = y["vid_indices"].to(self.device, non_blocking=True).long().reshape(B).contiguous()
        emo, emo_ptr, emo_stride = None, None, 0
        if self.dims.n_pre_emb == 2:
            emo = y["emo"].to(self.device, non_blocking=True).long()
            if emo.dim() < 1 or emo.shape[0] != B:
                raise LsError("y['emo'] must be [%d, ...] (got %s)" % (B, tuple(emo.shape)))
            emo_ptr, emo_stride = c_void_p(emo.data_ptr()), emo.stride(0)
        # nn.Embedding raises IndexError on an out-of-range index (RAG.py:116, scripts_beat/model/RAG.py:125); one
        # host sync per batch buys the same behaviour instead of a clamped gather
        lo, hi = int(vid.min()), int(vid.max())
        if lo < 0 or hi >= self.cfg.n_speakers:
            raise IndexError("y['vid_indices'] out of range [0,%d): min %d, max %d" % (self.cfg.n_speakers, lo, hi))
        if emo is not None:
            e0 = emo.reshape(B, -1)[:, 0]
            lo, hi = int(e0.min()), int(e0.max())
            if lo < 0 or hi >= self.cfg.n_emotions:
                raise IndexError("y['emo'] out of range [0,%d): min %d, max %d" % (self.cfg.n_emotions, lo, hi))
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_precompute_cond(self.h, B, c_void_p(audio.data_ptr()), c_void_p(ox.data_ptr()),
                                                    c_void_p(vid.data_ptr()), emo_ptr, emo_stride, 1, _stream()))
        if not inplace:
            ox_user[..., 4:] = 0          # keep the reference's visible side effect
        self._cond_keep = (audio, ox, vid, emo)
        self._cond_sig = self._sig(y)
        self.cond_batch = B

    def debug_buffer(self, which):
        """0 = A, 1 = P, 2 = z_mu, 3 = z_logvar, 4 = time-embedding table (flat fp32 copies)."""
        n = c_int64()
        self._check(self.lib.ls_debug_buffer(self.h, which, None, 0, byref(n), _stream()))
        out = torch.empty(n.value, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_debug_buffer(self.h, which, c_void_p(out.data_ptr()), n.value, byref(n), _stream()))
        return out

    def debug_hidden(self, layer, B):
        """Arm ls_debug_hidden: returns the [B,2,S,512] tensor the next denoiser launch fills with the residual stream
        after MLPblock `layer` (-1: after the input projection).  layer=None disarms."""
        if layer is None:
            self._check(self.lib.ls_debug_hidden(self.h, -1, None))
            self._dbg_keep = None
            return None
        S = 34 + self.dims.n_pre_emb
        out = torch.zeros(B, 2, S, self.dims.latent_dim, dtype=torch.float32, device=self.device)
        self._check(self.lib.ls_debug_hidden(self.h, int(layer), c_void_p(out.data_ptr())))
        self._dbg_keep = out
        return out

    # ---- compute ----------------------------------------------------------------------
    def wav_encoder(self, audio):
        audio = _f32(audio, self.device)
        out = torch.empty(audio.shape[0], 34, 256, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_wav_encoder(self.h, audio.shape[0], c_void_p(audio.data_ptr()),
                                                c_void_p(out.data_ptr()), _stream()))
        return out

    def _timesteps(self, t, B):
        """[B] int64 ORIGINAL timesteps on the device, checked against the time-embedding table (the reference indexes
        pe[timesteps], mlp_module.py:135, and raises on an out-of-range value; the kernels would clamp)."""
        t = t.to(self.device).long().contiguous()
        if t.numel() != B:
            raise LsError("timesteps must be [%d] (got %s)" % (B, tuple(t.shape)))
        lo, hi = int(t.min()), int(t.max())
        if lo < 0 or hi >= self.max_timestep:
            raise IndexError("timestep out of range [0,%d): min %d, max %d" % (self.max_timestep, lo, hi))
        return t

    def model_forward(self, x, t, uncond, style_eps):
        B = x.shape[0]
        x = _f32(x, self.device)
        t = self._timesteps(t, B)
        eps = _f32(style_eps, self.device)
        out = torch.empty(B, self.dims.njoints, self.dims.nfeats, 34, dtype=torch.float32, device=self.device)
        mu = torch.empty(B, 1, self.dims.latent_dim, dtype=torch.float32, device=self.device)
        lv = torch.empty_like(mu)
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_model_forward(self.h, B, c_void_p(x.data_ptr()), c_void_p(t.data_ptr()),
                                                  1 if uncond else 0, c_void_p(eps.data_ptr()),
                                                  c_void_p(out.data_ptr()), c_void_p(mu.data_ptr()),
                                                  c_void_p(lv.data_ptr()), _stream()))
        return _ref_layout(out), mu, lv

    def model_forward_train(self, x, t, cond_drop, style_eps):
        """ls_model_forward_train: RAG.forward with the training-mode per-clip condition dropout (forward only)."""
        B = x.shape[0]
        x = _f32(x, self.device)
        t = self._timesteps(t, B)
        eps = _f32(style_eps, self.device)
        drop = None if cond_drop is None else cond_drop.to(self.device).to(torch.uint8).contiguous()
        out = torch.empty(B, self.dims.njoints, self.dims.nfeats, 34, dtype=torch.float32, device=self.device)
        mu = torch.empty(B, 1, self.dims.latent_dim, dtype=torch.float32, device=self.device)
        lv = torch.empty_like(mu)
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_model_forward_train(self.h, B, c_void_p(x.data_ptr()), c_void_p(t.data_ptr()),
                                                        c_void_p(drop.data_ptr()) if drop is not None else None,
                                                        c_void_p(eps.data_ptr()), c_void_p(out.data_ptr()),
                                                        c_void_p(mu.data_ptr()), c_void_p(lv.data_ptr()), _stream()))
        return _ref_layout(out), mu, lv

    def cfg_forward(self, x, t, eps_c, eps_u, scale):
        B = x.shape[0]
        x = _f32(x, self.device)
        t = self._timesteps(t, B)
        eps_c, eps_u, scale = _f32(eps_c, self.device), _f32(eps_u, self.device), _f32(scale, self.device)
        out = torch.empty(B, self.dims.njoints, self.dims.nfeats, 34, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_cfg_forward(self.h, B, c_void_p(x.data_ptr()), c_void_p(t.data_ptr()),
                                                c_void_p(eps_c.data_ptr()), c_void_p(eps_u.data_ptr()),
                                                c_void_p(scale.data_ptr()), c_void_p(out.data_ptr()), _stream()))
        return _ref_layout(out)

    def cfg_forward_grad(self, x, t, eps_c, eps_u, scale):
        """ls_cfg_forward_grad: cfg_forward on the exact-order fp32 path, keeping what ls_cfg_backward needs."""
        B = x.shape[0]
        x = _f32(x, self.device)
        t = self._timesteps(t, B)
        eps_c, eps_u, scale = _f32(eps_c, self.device), _f32(eps_u, self.device), _f32(scale, self.device)
        out = torch.empty(B, self.dims.njoints, self.dims.nfeats, 34, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_cfg_forward_grad(self.h, B, c_void_p(x.data_ptr()), c_void_p(t.data_ptr()),
                                                     c_void_p(eps_c.data_ptr()), c_void_p(eps_u.data_ptr()),
                                                     c_void_p(scale.data_ptr()), c_void_p(out.data_ptr()), _stream()))
        return _ref_layout(out)

    def cfg_backward(self, grad_out, scale):
        """ls_cfg_backward: J^T grad_out for the last cfg_forward_grad call (J = d out / d x)."""
        B = grad_out.shape[0]
        g = _f32(grad_out, self.device)
        scale = _f32(scale, self.device)
        gx = torch.empty(B, self.dims.njoints, self.dims.nfeats, 34, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_cfg_backward(self.h, B, c_void_p(g.data_ptr()), c_void_p(scale.data_ptr()),
                                                 c_void_p(gx.data_ptr()), _stream()))
        return gx

    def step(self, params, x_t, eps_c, eps_u, noise, scale, x_prev, pred_x0):
        """One fused denoising step.  x_t / x_prev / pred_x0: dense fp32 [B,J,D,F] on the
        device; noise: any [B,J,D,F] tensor whose (J,D) dims collapse, else it is copied."""
        B = x_t.shape[0]
        nb = nj = nf = 0
        nptr = None
        if noise is not None:
            sb, sj, sd, sf = noise.stride()
            if noise.dtype != torch.float32 or noise.device != self.device or sj != sd * noise.shape[2]:
                noise = _f32(noise, self.device)
                sb, sj, sd, sf = noise.stride()
            nb, nj, nf, nptr = sb, sd, sf, c_void_p(noise.data_ptr())
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_step(self.h, B, byref(params), c_void_p(x_t.data_ptr()),
                                         c_void_p(eps_c.data_ptr()), c_void_p(eps_u.data_ptr()), nptr, nb, nj, nf,
                                         c_void_p(scale.data_ptr()),
                                         c_void_p(x_prev.data_ptr()) if x_prev is not None else None,
                                         c_void_p(pred_x0.data_ptr()) if pred_x0 is not None else None, _stream()))

    def _noise_arg(self, noise):
        """(tensor kept alive, ptr, sb, sj, sf) of a step-noise tensor whose (J,D) dims collapse."""
        if noise is None:
            return None, None, 0, 0, 0
        sb, sj, sd, sf = noise.stride()
        if noise.dtype != torch.float32 or noise.device != self.device or sj != sd * noise.shape[2]:
            noise = _f32(noise, self.device)
            sb, sj, sd, sf = noise.stride()
        return noise, noise.data_ptr(), sb, sd, sf

    def graphed_draws(self, K, B, d, perm_like, factory):
        """Per-engine cache of the CUDA-graph-captured draws of a chunk (gaussian_diffusion._GraphedDraws)."""
        key = (factory.__name__, K, B, d, tuple(perm_like.shape), tuple(perm_like.stride()))
        cache = self.__dict__.setdefault("_graphed", {})
        if key not in cache:
            cache[key] = factory(K, B, d, perm_like)
        return cache[key]

    def step_multi(self, params, x_t, eps_c, eps_u, noise, scale, x_prev, pred_x0):
        """len(params) <= MAX_FUSED_STEPS consecutive steps in one launch (ls_step_multi).
        eps_c / eps_u / noise: per-step lists; x_prev / pred_x0: dense [K,B,J,D,F] outputs
        (pred_x0 may be None).  Step k starts from x_t (k = 0) or x_prev[k-1]."""
        K = len(params)
        B = x_t.shape[0]
        P = (LsStepParams * K)(*params)
        IO = (LsStepIO * K)()
        keep = []
        for k in range(K):
            nz, nptr, sb, sj, sf = self._noise_arg(noise[k])
            keep.append(nz)
            IO[k].eps_cond, IO[k].eps_uncond = eps_c[k].data_ptr(), eps_u[k].data_ptr()
            IO[k].noise, IO[k].noise_sb, IO[k].noise_sj, IO[k].noise_sf = nptr, sb, sj, sf
            IO[k].x_prev = x_prev[k].data_ptr()
            IO[k].pred_x0 = pred_x0[k].data_ptr() if pred_x0 is not None else None
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_step_multi(self.h, B, K, P, IO, c_void_p(x_t.data_ptr()),
                                               c_void_p(scale.data_ptr()), _stream()))

    def q_sample(self, x0, noise, c_x0, c_noise):
        x0 = _f32(x0, self.device)
        noise = _f32(noise, self.device)
        out = torch.empty_like(x0)
        with torch.cuda.device(self.device):
            self._check(self.lib.ls_q_sample(self.h, x0.numel(), c_void_p(x0.data_ptr()), c_void_p(noise.data_ptr()),
                                             float(c_x0), float(c_noise), c_void_p(out.data_ptr()), _stream()))
        return out


def huber_terms(target, output, z_mu, z_logvar, terms):
    """ls_huber_terms: {rot_mse, vel_mse, kld} of training_losses' HUBER branch into terms (3 floats on the device)."""
    lib = load_library()
    lib.ls_huber_terms.argtypes = [c_int64, c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int32,
                                   c_void_p]
    F = target.shape[-1]
    rows = target.numel() // F
    with torch.cuda.device(target.device):
        rc = lib.ls_huber_terms(rows, F, c_void_p(target.data_ptr()), c_void_p(output.data_ptr()),
                                0 if z_mu is None else z_mu.numel(),
                                None if z_mu is None else c_void_p(z_mu.data_ptr()),
                                None if z_logvar is None else c_void_p(z_logvar.data_ptr()),
                                c_void_p(terms.data_ptr()), target.device.index or 0, _stream())
    if rc != 0:
        raise LsError("libls_b200 error %d: %s" % (rc, lib.ls_last_error(None).decode()))


def vb_terms(x_start, mean1, mean2, logvar1, logvar2, t):
    """ls_vb_terms: per-clip variational-bound term in bits.  x_start / mean1 / mean2 [B, ...] (mean2 None = 0),
    logvar1 / logvar2 [B] (logvar2 None = 0), t [B] int64 or None (None: always the KL)."""
    if mean1.device.type != "cuda":
        raise LsError("the variational-bound terms run on a CUDA sm_100 device only (no CPU path)")
    lib = load_library()
    lib.ls_vb_terms.argtypes = [c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int32, c_void_p]
    B = mean1.shape[0]
    dev = mean1.device

    def dense(v, dtype=torch.float32):
        return None if v is None else v.detach().to(device=dev, dtype=dtype).contiguous()

    xs, m1, m2 = dense(x_start), dense(mean1), dense(mean2)
    lv1, lv2, tt = dense(logvar1), dense(logvar2), dense(t, torch.int64)
    out = torch.empty(B, dtype=torch.float32, device=dev)
    ptr = lambda v: None if v is None else c_void_p(v.data_ptr())      # noqa: E731
    with torch.cuda.device(dev):
        rc = lib.ls_vb_terms(B, m1.numel() // B, ptr(xs), ptr(m1), ptr(m2), ptr(lv1), ptr(lv2), ptr(tt), ptr(out),
                             dev.index or 0, _stream())
    if rc != 0:
        raise LsError("libls_b200 error %d: %s" % (rc, lib.ls_last_error(None).decode()))
    return out
