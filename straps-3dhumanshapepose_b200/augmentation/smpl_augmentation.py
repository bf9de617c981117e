"""SMPL shape / pose augmentation on the GPU (drop-in for reference augmentation/smpl_augmentation.py:6-61).

The random draws come from torch's generator with the same calls, shapes and order as the reference (so the same seed
gives the same stream on the same device); the affine part and the axis-angle -> rotation-matrix conversion are library
kernels (csrc/synth.cu).  Given the same draws the sampled shapes are bit-identical to the reference's.
"""
import torch

from straps_b200 import ops


def uniform_sample_shape(batch_size, mean_shape, delta_betas_range):
    """mean_shape + U(l, h) per coefficient -> (bs, 10)   (reference lines 6-14)."""
    l, h = delta_betas_range
    noise = torch.rand(batch_size, 10, device=mean_shape.device)
    return ops.scale_shift(noise, h - l, l, mean_shape)


def normal_sample_shape(batch_size, mean_shape, std_vector):
    """mean_shape + N(0, std_vector^2) -> (bs, 10)   (reference lines 17-24)."""
    noise = torch.randn(batch_size, 10, device=mean_shape.device)
    return ops.scale_shift(noise, std_vector, None, mean_shape)


def augment_smpl(orig_shape, pose, global_orients, mean_shape, smpl_augment_params):
    """-> (new_shape [bs,10], pose_rotmats [bs,23,3,3], glob_rotmats [bs,1,3,3])   (reference lines 27-61)."""
    batch_size = orig_shape.shape[0]
    if smpl_augment_params['augment_shape']:
        distribution = smpl_augment_params['delta_betas_distribution']
        assert distribution in ['uniform', 'normal']
        if distribution == 'uniform':
            new_shape = uniform_sample_shape(batch_size, mean_shape, smpl_augment_params['delta_betas_range'])
        else:
            assert smpl_augment_params['delta_betas_std_vector'] is not None
            new_shape = normal_sample_shape(batch_size, mean_shape, smpl_augment_params['delta_betas_std_vector'])
    else:
        new_shape = orig_shape
    # one launch for the 24 joints of every body, split afterwards
    rotmats = ops.batch_rodrigues(torch.cat([global_orients.reshape(-1, 1, 3), pose.reshape(-1, 23, 3)], dim=1).reshape(-1, 3))
    rotmats = rotmats.view(-1, 24, 3, 3)
    return new_shape, rotmats[:, 1:], rotmats[:, :1]
