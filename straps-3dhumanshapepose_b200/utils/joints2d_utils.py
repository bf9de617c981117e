"""2-D joint helpers used around the hot path (reference utils/joints2d_utils.py)."""
import torch


def undo_keypoint_normalisation(normalised_keypoints, img_wh):
    """[-1, 1] -> pixel coordinates (reference lines 5-10)."""
    return (normalised_keypoints + 1) * (img_wh / 2.0)


def check_joints2d_visibility_torch(joints2d, img_wh):
    """Bool [B,N]: joint strictly inside [0, img_wh] on both axes (reference lines 23-33)."""
    x, y = joints2d[:, :, 0], joints2d[:, :, 1]
    return ~((x > img_wh) | (y > img_wh) | (x < 0) | (y < 0))
