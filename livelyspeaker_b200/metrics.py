"""Post-sampling rhythm metric of the TED evaluation on the device.

Mirrors the body of ``infer_from_testloader`` between the sampler call and the score accumulation
(scripts/test_RAG_ted.py:84-123): direction vectors -> joint-pair angles -> per-frame angle change -> motion beats ->
beat alignment score against audio onset times.  The arithmetic runs in two small CUDA kernels behind the C ABI
(``ls_motion_beats`` / ``ls_beat_align``) on the sampler's output where it lies; there is no CPU path (``LsError``).
The onset detector (``librosa.onset.onset_detect``, :113) stays with the caller: pass its result as ``audio_beat_times``.
The FGD / feature-distance / diversity scores of the same script live in ``ted_evaluator.py`` / ``embedding_net.py``.
"""
import ctypes
from ctypes import c_float, c_int32, c_void_p

import torch

from . import _cabi

# constants of the evaluation script (scripts/test_RAG_ted.py:22-34)
MEAN_DIR_VEC = [0.0154009, -0.9690125, -0.0884354, -0.0022264, -0.8655276, 0.4342174, -0.0035145, -0.8755367,
                -0.4121039, -0.9236511, 0.3061306, -0.0012415, -0.5155854, 0.8129665, 0.0871897, 0.2348464,
                0.1846561, 0.8091402, 0.9271948, 0.2960011, -0.013189, 0.5233978, 0.8092403, 0.0725451, -0.2037076,
                0.1924306, 0.8196916]
ANGLE_PAIR = [(3, 4), (4, 5), (6, 7), (7, 8)]
CHANGE_ANGLE = [0.0034540758933871984, 0.007043459918349981, 0.003493624273687601, 0.007205077446997166]
THRES = 0.03
SIGMA = 0.1
FPS = 15.0


def _dev_index(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def motion_beats(sample, mean_dir_vec=MEAN_DIR_VEC, angle_pair=ANGLE_PAIR, change_angle=CHANGE_ANGLE, thres=THRES):
    """sample: the sampler's output [B, njoints, 3, F] on a CUDA device -> (angle_diff [B,F] fp32, beat_mask [B,F] uint8).
    scripts/test_RAG_ted.py:84-111."""
    if sample.device.type != "cuda":
        raise _cabi.LsError("metrics.motion_beats runs on a CUDA sm_100 device only (no CPU path)")
    B, J, D, F = sample.shape
    if D != 3:
        raise ValueError("direction vectors have 3 components (got nfeats=%d)" % D)
    if len(mean_dir_vec) != J * 3 or len(angle_pair) != len(change_angle):
        raise ValueError("mean_dir_vec must have %d entries and angle_pair / change_angle the same length" % (J * 3))
    lib = _cabi.load_library()
    x = sample.float().contiguous()
    angle_diff = torch.empty(B, F, dtype=torch.float32, device=x.device)
    mask = torch.empty(B, F, dtype=torch.uint8, device=x.device)
    n = len(angle_pair)
    mean = (c_float * (J * 3))(*[float(v) for v in mean_dir_vec])
    pairs = (c_int32 * (2 * n))(*[int(j) for p in angle_pair for j in p])
    change = (c_float * n)(*[float(v) for v in change_angle])
    with torch.cuda.device(x.device):
        rc = lib.ls_motion_beats(B, J, F, c_void_p(x.data_ptr()), mean, pairs, change, n, float(thres),
                                 c_void_p(angle_diff.data_ptr()), c_void_p(mask.data_ptr()), _dev_index(x),
                                 c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise _cabi.LsError("libls_b200 error %d: %s" % (rc, lib.ls_last_error(None).decode()))
    return angle_diff, mask


def beat_align_score(beat_mask, audio_beat_times, sigma=SIGMA, fps=FPS):
    """beat_mask [B,F] (from motion_beats); audio_beat_times: per clip a sequence of onset times in seconds.
    Returns the three accumulators of the script for this batch plus the per-clip values
    (scripts/test_RAG_ted.py:112-123): beat_align_score_sum, num_beats, motion_beats_sum."""
    if beat_mask.device.type != "cuda":
        raise _cabi.LsError("metrics.beat_align_score runs on a CUDA sm_100 device only (no CPU path)")
    B, F = beat_mask.shape
    if len(audio_beat_times) != B:
        raise ValueError("one onset list per clip")
    M = max(1, max(len(a) for a in audio_beat_times))
    ab = torch.zeros(B, M, dtype=torch.float32)
    na = torch.zeros(B, dtype=torch.int32)
    for b, a in enumerate(audio_beat_times):
        if len(a):
            ab[b, :len(a)] = torch.as_tensor(list(a), dtype=torch.float32)
        na[b] = len(a)
    dev = beat_mask.device
    ab, na = ab.to(dev), na.to(dev)
    mask = beat_mask.to(torch.uint8).contiguous()
    score = torch.empty(B, dtype=torch.float64, device=dev)
    n_motion = torch.empty(B, dtype=torch.int32, device=dev)
    n_audio = torch.empty(B, dtype=torch.int32, device=dev)
    lib = _cabi.load_library()
    with torch.cuda.device(dev):
        rc = lib.ls_beat_align(B, F, c_void_p(mask.data_ptr()), c_void_p(ab.data_ptr()), c_void_p(na.data_ptr()), M,
                               float(fps), float(sigma), c_void_p(score.data_ptr()), c_void_p(n_motion.data_ptr()),
                               c_void_p(n_audio.data_ptr()), _dev_index(mask),
                               c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise _cabi.LsError("libls_b200 error %d: %s" % (rc, lib.ls_last_error(None).decode()))
    return {"beat_align_score_sum": float(score.sum()), "num_beats": int(n_audio.sum()),
            "motion_beats_sum": int(n_motion.sum()), "clip_score": score, "clip_n_motion": n_motion,
            "clip_n_audio": n_audio}
