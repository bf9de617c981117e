"""ORACLE (test infrastructure): restatement of smplx.lbs."""
import torch
import torch.nn.functional as F


def blend_shapes(betas, shape_disps):
    # S2: [B,L] x [V,3,L] -> [B,V,3]
    return torch.einsum('bl,mkl->bmk', [betas, shape_disps])


def vertices2joints(J_regressor, vertices):
    # S3: [J,V] x [B,V,3] -> [B,J,3]
    return torch.einsum('bik,ji->bjk', [vertices, J_regressor])


def batch_rodrigues(rot_vecs, epsilon=1e-8, dtype=torch.float32):
    # Appendix A: the epsilon is added INSIDE the norm
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    axis = rot_vecs / angle
    c = torch.cos(angle).unsqueeze(1)
    s = torch.sin(angle).unsqueeze(1)
    rx, ry, rz = torch.split(axis, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype, device=rot_vecs.device)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device).unsqueeze(0)
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def transform_mat(R, t):
    # [N,3,3],[N,3,1] -> [N,4,4]
    return torch.cat([F.pad(R, [0, 0, 0, 1]), F.pad(t, [0, 0, 0, 1], value=1)], dim=2)


def batch_rigid_transform(rot_mats, joints, parents, dtype=torch.float32):
    # S5
    joints = torch.unsqueeze(joints, dim=-1)
    rel = joints.clone()
    rel[:, 1:] -= joints[:, parents[1:]]
    T = transform_mat(rot_mats.reshape(-1, 3, 3), rel.reshape(-1, 3, 1)).reshape(-1, joints.shape[1], 4, 4)
    chain = [T[:, 0]]
    for i in range(1, parents.shape[0]):
        chain.append(torch.matmul(chain[int(parents[i])], T[:, i]))
    G = torch.stack(chain, dim=1)
    posed_joints = G[:, :, :3, 3]
    joints_h = F.pad(joints, [0, 0, 0, 1])
    A = G - F.pad(torch.matmul(G, joints_h), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, A


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights,
        pose2rot=True, dtype=torch.float32):
    B = max(betas.shape[0], pose.shape[0])
    device = betas.device
    v_shaped = v_template + blend_shapes(betas, shapedirs)
    J = vertices2joints(J_regressor, v_shaped)
    ident = torch.eye(3, dtype=dtype, device=device)
    if pose2rot:
        rot_mats = batch_rodrigues(pose.view(-1, 3), dtype=dtype).view([B, -1, 3, 3])
        pose_feature = (rot_mats[:, 1:, :, :] - ident).view([B, -1])
        pose_offsets = torch.matmul(pose_feature, posedirs).view(B, -1, 3)
    else:
        pose_feature = pose[:, 1:].view(B, -1, 3, 3) - ident
        rot_mats = pose.view(B, -1, 3, 3)
        pose_offsets = torch.matmul(pose_feature.view(B, -1), posedirs).view(B, -1, 3)
    v_posed = pose_offsets + v_shaped
    J_transformed, A = batch_rigid_transform(rot_mats, J, parents, dtype=dtype)
    W = lbs_weights.unsqueeze(dim=0).expand([B, -1, -1])
    nj = J_regressor.shape[0]
    T = torch.matmul(W, A.view(B, nj, 16)).view(B, -1, 4, 4)
    ones = torch.ones([B, v_posed.shape[1], 1], dtype=dtype, device=device)
    v_h = torch.matmul(T, torch.unsqueeze(torch.cat([v_posed, ones], dim=2), dim=-1))
    return v_h[:, :, :3, 0], J_transformed
