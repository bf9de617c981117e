"""torch.autograd.Function wrappers: PyTorch's engine chains the library's forward/backward kernels."""
import torch

from . import ops


class Rot6dToRotmat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x6 = x.contiguous().view(-1, 6)
        ctx.save_for_backward(x6)
        ctx.in_shape = x.shape
        return ops.rot6d_to_rotmat(x6)

    @staticmethod
    def backward(ctx, g):
        (x6,) = ctx.saved_tensors
        return ops.rot6d_backward(x6, g).view(ctx.in_shape)


class BatchRodrigues(torch.autograd.Function):
    """axis-angle [n,3] -> rotation matrices [n,3,3] (smplx.lbs.batch_rodrigues), differentiable."""

    @staticmethod
    def forward(ctx, rot_vecs):
        r = rot_vecs.contiguous()
        ctx.save_for_backward(r)
        return ops.batch_rodrigues(r)

    @staticmethod
    def backward(ctx, g):
        (r,) = ctx.saved_tensors
        return ops.batch_rodrigues_backward(r, g)


class OrthographicProject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points3d, cam):
        ctx.save_for_backward(points3d, cam)
        return ops.orthographic_project(points3d, cam)

    @staticmethod
    def backward(ctx, g):
        points3d, cam = ctx.saved_tensors
        dp, dc = ops.orthographic_project_backward(points3d, cam, g)
        return dp, dc


class SmplForward(torch.autograd.Function):
    """(rotmats [B,24,3,3], betas [B,10]) -> (vertices [B,6890,3], joints [B,90,3])."""

    @staticmethod
    def forward(ctx, handle, rotmats, betas):
        verts, joints, saved = handle.forward_train(rotmats, betas)
        ctx.handle = handle
        ctx.save_for_backward(*saved)
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, g_joints):
        d_rot, d_betas = ctx.handle.backward(ctx.saved_tensors, g_verts, g_joints)
        return None, d_rot, d_betas
