"""Camera-translation augmentation on the GPU (drop-in for reference augmentation/cam_augmentation.py:4-14)."""
import torch

from straps_b200 import ops


def augment_cam_t(mean_cam_t, xy_std=0.05, delta_z_range=[-5, 5]):
    """mean_cam_t [bs,3] -> [bs,3]: x, y += N(0, xy_std^2); z += U(l, h).  Draw order as the reference: randn(bs,2), rand(bs)."""
    batch_size = mean_cam_t.shape[0]
    device = mean_cam_t.device
    l, h = delta_z_range
    noise = torch.cat([torch.randn(batch_size, 2, device=device), torch.rand(batch_size, device=device).unsqueeze(1)], dim=1)
    mul = torch.tensor([xy_std, xy_std, h - l], dtype=torch.float32, device=device)
    add = torch.tensor([0.0, 0.0, l], dtype=torch.float32, device=device)
    return ops.scale_shift(noise, mul, add, mean_cam_t)
