"""6-D rotation decode on the GPU (drop-in for reference utils/rigid_transform_utils.py:27-41)."""
import torch

from straps_b200 import ops
from straps_b200.autograd import Rot6dToRotmat


def rot6d_to_rotmat(x):
    """(B, 6k) interleaved 6-D rotations (a1 = elements 0,2,4; a2 = 1,3,5) -> (B*k, 3, 3), columns (b1, b2, b3)."""
    if torch.is_grad_enabled() and x.requires_grad:
        return Rot6dToRotmat.apply(x)
    return ops.rot6d_to_rotmat(x)
