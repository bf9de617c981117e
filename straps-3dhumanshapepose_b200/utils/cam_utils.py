"""Weak-perspective projection on the GPU (drop-in for reference utils/cam_utils.py:5-26)."""
import torch

from straps_b200 import ops
from straps_b200.autograd import OrthographicProject


def orthographic_project_torch(points3D, cam_params):
    """points3D [B,N,3], cam_params [B,3] = (s, tx, ty)  ->  [B,N,2] = s * (xy + t)."""
    if torch.is_grad_enabled() and (points3D.requires_grad or cam_params.requires_grad):
        return OrthographicProject.apply(points3D.contiguous(), cam_params.contiguous())
    return ops.orthographic_project(points3D, cam_params)
