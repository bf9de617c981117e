"""CPU: mathematical invariants of the SMPL restatement -- the SMPL arithmetic lives in
the absent third-party smplx package, so these properties stand in for golden vectors the reference lacks."""
import numpy as np
import torch

import straps_oracle as O
from conftest import rel_err


def test_identity_pose_reduces_to_shape_blend(smpl_oracle):
    B = 3
    betas = torch.from_numpy(np.random.RandomState(0).normal(0, 1, (B, 10)).astype(np.float32))
    R = torch.eye(3).repeat(B, 24, 1, 1)
    with torch.no_grad():
        v, j = smpl_oracle.forward_rotmats(R, betas)
        m = smpl_oracle.smpl
        v_shaped = m.v_template + torch.einsum('bl,mkl->bmk', betas, m.shapedirs)
        J = torch.einsum('bik,ji->bjk', v_shaped, m.J_regressor)
    assert rel_err(v.numpy(), v_shaped.numpy()) < 1e-6
    assert rel_err(j[:, :24].numpy(), J.numpy()) < 1e-6


def test_root_rotation_is_rigid(smpl_oracle):
    B = 2
    rng = np.random.RandomState(1)
    betas = torch.from_numpy(rng.normal(0, 1, (B, 10)).astype(np.float32))
    R0 = O.rot6d_to_rotmat(torch.from_numpy(rng.normal(0, 1, (B, 6)).astype(np.float32)))
    R = torch.eye(3).repeat(B, 24, 1, 1)
    R[:, 0] = R0
    with torch.no_grad():
        v, j = smpl_oracle.forward_rotmats(R, betas)
        m = smpl_oracle.smpl
        v_shaped = m.v_template + torch.einsum('bl,mkl->bmk', betas, m.shapedirs)
        J0 = torch.einsum('bik,ji->bjk', v_shaped, m.J_regressor)[:, 0:1]
        expect = torch.einsum('bij,bvj->bvi', R0, v_shaped - J0) + J0
    assert rel_err(v.numpy(), expect.numpy()) < 1e-5


def test_rot6d_is_orthonormal_with_det_one():
    x = torch.from_numpy(np.random.RandomState(2).normal(0, 1, (64, 6)).astype(np.float32))
    R = O.rot6d_to_rotmat(x)
    eye = torch.eye(3).expand_as(R)
    assert (R.transpose(1, 2) @ R - eye).abs().max() < 1e-5
    assert (torch.linalg.det(R) - 1).abs().max() < 1e-5
    # interleaved layout: a1 = elements (0,2,4)
    a1 = x[:, [0, 2, 4]]
    assert torch.allclose(R[:, :, 0], a1 / a1.norm(dim=1, keepdim=True), atol=1e-6)


def test_skinning_weights_partition_of_unity_and_picks(smpl_oracle):
    m = smpl_oracle.smpl
    assert torch.allclose(m.lbs_weights.sum(1), torch.ones(6890), atol=1e-6)
    assert int((m.lbs_weights != 0).sum(1).max()) <= 4
    betas = torch.zeros(1, 10)
    R = O.rot6d_to_rotmat(torch.from_numpy(np.random.RandomState(3).normal(0, 1, (1, 144)).astype(np.float32))).view(1, 24, 3, 3)
    with torch.no_grad():
        v, j = smpl_oracle.forward_rotmats(R, betas)
    idx = m.vertex_joint_selector.extra_joints_idxs
    assert torch.equal(j[:, 24:45], v[:, idx])          # bit exact vertex picks
    assert idx.tolist()[:5] == [332, 6260, 2800, 4071, 583]
