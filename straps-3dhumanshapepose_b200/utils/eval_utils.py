"""Mesh / skeleton alignment for the evaluation metrics, on the GPU (drop-in for reference utils/eval_utils.py).

The reference aligns one sample at a time in numpy (`np.linalg.svd` inside a Python loop over the batch); here a batch is
one kernel launch (csrc/metrics.cu: one CTA per body, fp64 3x3 solve).  Inputs may be CUDA tensors (returned as CUDA
tensors, no host round trip) or numpy arrays as in the reference (uploaded, aligned on the device, returned as numpy):
either way the arithmetic runs on the GPU.
"""
import numpy as np
import torch

from straps_b200 import ops
from straps_b200._lib import StrapsError


def _to_device(a, name):
    if torch.is_tensor(a):
        if not a.is_cuda:
            raise StrapsError('%s: CPU tensors are not accepted (pass CUDA tensors or numpy arrays)' % name)
        return a.float(), None
    if not torch.cuda.is_available():
        raise StrapsError('%s needs a CUDA device: the B200 path has no CPU fallback' % name)
    arr = np.asarray(a)
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).cuda(), arr.dtype


def _back(t, np_dtype):
    return t if np_dtype is None else t.cpu().numpy().astype(np_dtype, copy=False)


def procrustes_analysis_batch(S1, S2):
    """[B,N,3] x2 -> S1 after the similarity transform (scale, rotation, translation) closest to S2 (reference lines 55-60)."""
    a, dt = _to_device(S1, 'procrustes_analysis_batch')
    b, _ = _to_device(S2, 'procrustes_analysis_batch')
    _, out = ops.points_metrics(a, b, ops.METRIC_PA, want_pa=True)
    return _back(out, dt)


def compute_similarity_transform(S1, S2):
    """Single-sample form of the above (reference lines 7-52): [N,3] (or [3,N]) point sets."""
    a, dt = _to_device(S1, 'compute_similarity_transform')
    b, _ = _to_device(S2, 'compute_similarity_transform')
    # the reference works on [3,N] and transposes anything whose leading axis is not 3 (or 2); the kernel works on [N,3]
    coords_first = a.shape[0] in (2, 3)
    if coords_first:
        a, b = a.t(), b.t()
    if a.shape[1] != 3:
        raise StrapsError('compute_similarity_transform: only 3-D point sets are built')
    _, out = ops.points_metrics(a.contiguous()[None], b.contiguous()[None], ops.METRIC_PA, want_pa=True)
    out = out[0].t() if coords_first else out[0]
    return _back(out, dt)


def scale_and_translation_transform_batch(P, T):
    """[B,N,3] x2 -> P moved to T's mean and RMS distance from the mean (reference lines 63-85)."""
    a, dt = _to_device(P, 'scale_and_translation_transform_batch')
    b, _ = _to_device(T, 'scale_and_translation_transform_batch')
    out, _ = ops.points_metrics(a, b, ops.METRIC_SC, want_sc=True)
    return _back(out, dt)
