"""Weak-perspective projection on the GPU (drop-in for reference utils/cam_utils.py:5-26)."""
from straps_b200 import ops


def orthographic_project_torch(points3D, cam_params):
    """points3D [B,N,3], cam_params [B,3] = (s, tx, ty)  ->  [B,N,2] = s * (xy + t)."""
    return ops.orthographic_project(points3D, cam_params)
