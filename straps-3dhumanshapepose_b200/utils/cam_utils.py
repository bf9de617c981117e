"""Camera projections on the GPU (drop-in for reference utils/cam_utils.py)."""
import numpy as np
import torch

from straps_b200 import ops
from straps_b200._lib import StrapsError
from straps_b200.autograd import OrthographicProject


def orthographic_project_torch(points3D, cam_params):
    """points3D [B,N,3], cam_params [B,3] = (s, tx, ty)  ->  [B,N,2] = s * (xy + t)   (reference lines 5-26)."""
    if torch.is_grad_enabled() and (points3D.requires_grad or cam_params.requires_grad):
        return OrthographicProject.apply(points3D.contiguous(), cam_params.contiguous())
    return ops.orthographic_project(points3D, cam_params)


def get_intrinsics_matrix(img_width, img_height, focal_length):
    """3x3 calibration matrix with the principal point at the image centre (reference lines 29-38)."""
    K = np.zeros((3, 3))
    K[0, 0] = K[1, 1] = focal_length
    K[0, 2], K[1, 2], K[2, 2] = img_width / 2.0, img_height / 2.0, 1.
    return K


def perspective_project_torch(points, rotation, translation, cam_K=None, focal_length=None, img_wh=None):
    """points [bs,N,3], rotation [bs,3,3], translation [bs,3], cam_K [bs,3,3] (or focal_length + img_wh) -> [bs,N,2]
    (reference lines 40-71).  One kernel: rotate + translate, divide by depth, apply the intrinsics."""
    batch_size = points.shape[0]
    if cam_K is None:
        K = torch.from_numpy(get_intrinsics_matrix(img_wh, img_wh, focal_length).astype(np.float32))
        cam_K = K[None].expand(batch_size, -1, -1).to(points.device)
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in (points, rotation, translation, cam_K)):
        # the reference's version is differentiable but only ever runs under torch.no_grad() (train/...:112,141-143); returning a
        # detached result here would silently zero a projection loss, so say so instead
        raise StrapsError('perspective_project_torch: the B200 kernel has no backward (the reference only calls it under '
                          'torch.no_grad() on the target side); detach the inputs or wrap the call in torch.no_grad()')
    return ops.perspective_project(points, rotation, translation, cam_K)


def convert_weak_perspective_to_camera_translation(cam_wp, focal_length, resolution):
    """(s, tx, ty) -> (tx, ty, 2f / (res * s + 1e-9))   (reference lines 74-76; host-side, B=1 pre-processing)."""
    return np.array([cam_wp[1], cam_wp[2], 2 * focal_length / (resolution * cam_wp[0] + 1e-9)])


def batch_convert_weak_perspective_to_camera_translation(wp_cams, focal_length, resolution):
    """[n,3] weak-perspective cameras -> [n,3] float32 translations (reference lines 79-87)."""
    wp = np.asarray(wp_cams)
    out = np.stack([wp[:, 1], wp[:, 2], 2 * focal_length / (resolution * wp[:, 0] + 1e-9)], axis=1)
    return out.astype(np.float32)
