"""Proxy-representation synthesis on the GPU.

Drop-in for the two functions of the reference's utils/label_conversions.py that feed the regressor
(train/train_synthetic_otf_rendering.py:178-182): the binary silhouette and the 2-D joint Gaussian heat-maps.  The
reference builds the heat-maps with a B x J Python loop and ~10 `.item()` device syncs per joint; here it is one memset and
one paste kernel (csrc/smpl.cu).  Results are bit-identical: the 16x16 window is computed once on the host with the very
same torch ops the reference uses (linspace / sqrt / pow / exp) and only copied by the kernel.
"""
import ctypes

import torch

from straps_b200 import _lib
from straps_b200._lib import StrapsError, check

_TABLES = {}


def _gaussian_table(std, device):
    key = (int(std), str(device))
    if key not in _TABLES:
        size = 2 * std                                                       # reference line 101
        x, y = torch.meshgrid(torch.linspace(-size, size, 2 * size), torch.linspace(-size, size, 2 * size), indexing='ij')
        d = torch.sqrt(x * x + y * y)
        _TABLES[key] = torch.exp(-(d ** 2 / (2.0 * std ** 2))).contiguous().to(device)
    return _TABLES[key]


def _cuda_f32(t, name):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise StrapsError('%s must be a CUDA tensor: the B200 path has no CPU fallback' % name)
    return t.float().contiguous()


def convert_multiclass_to_binary_labels_torch(multiclass_labels):
    """1 where the part label is non-zero, 0 elsewhere (same dtype/shape) -- reference lines 48-55."""
    x = _cuda_f32(multiclass_labels, 'multiclass_labels')
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.lib().straps_multiclass_to_binary(ctypes.c_void_p(x.data_ptr()), x.numel(), ctypes.c_void_p(out.data_ptr()),
                                                     ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
              'straps_multiclass_to_binary')
    return out.to(multiclass_labels.dtype)


def convert_2Djoints_to_gaussian_heatmaps_torch(joints2D, img_wh, std=4):
    """(B, N, 2) pixel joints -> (B, N, img_wh, img_wh) heat-maps, Gaussian truncated at 2 std -- reference lines 90-127."""
    j = _cuda_f32(joints2D, 'joints2D')
    B, N = j.shape[0], j.shape[1]
    table = _gaussian_table(std, j.device)
    out = torch.empty((B, N, img_wh, img_wh), dtype=torch.float32, device=j.device)
    with torch.cuda.device(j.device):
        check(_lib.lib().straps_joints2d_to_heatmaps(ctypes.c_void_p(j.data_ptr()), B, N, img_wh, 2 * std, ctypes.c_void_p(table.data_ptr()),
                                                     ctypes.c_void_p(out.data_ptr()),
                                                     ctypes.c_void_p(torch.cuda.current_stream(j.device).cuda_stream)),
              'straps_joints2d_to_heatmaps')
    return out
