"""ORACLE (test infrastructure): restatement of smplx.body_models.SMPL."""
import os
import pickle
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

from .lbs import lbs
from .vertex_ids import vertex_ids as VERTEX_IDS

ModelOutput = namedtuple('ModelOutput',
                         ['vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose',
                          'expression', 'left_hand_pose', 'right_hand_pose', 'jaw_pose'])
ModelOutput.__new__.__defaults__ = (None,) * len(ModelOutput._fields)


def _np(a, dtype=np.float32):
    if 'scipy.sparse' in str(type(a)):
        a = a.todense()
    return np.array(a, dtype=dtype)


class VertexJointSelector(nn.Module):
    def __init__(self, vertex_ids, use_hands=True, use_feet_keypoints=True):
        super().__init__()
        idx = [vertex_ids[k] for k in ('nose', 'reye', 'leye', 'rear', 'lear')]
        if use_feet_keypoints:
            idx += [vertex_ids[k] for k in ('LBigToe', 'LSmallToe', 'LHeel', 'RBigToe', 'RSmallToe', 'RHeel')]
        if use_hands:
            for side in 'lr':
                idx += [vertex_ids[side + t] for t in ('thumb', 'index', 'middle', 'ring', 'pinky')]
        self.register_buffer('extra_joints_idxs', torch.tensor(idx, dtype=torch.long))

    def forward(self, vertices, joints):
        return torch.cat([joints, torch.index_select(vertices, 1, self.extra_joints_idxs)], dim=1)


class SMPL(nn.Module):
    NUM_JOINTS = 23
    NUM_BODY_JOINTS = 23
    NUM_BETAS = 10

    def __init__(self, model_path, data_struct=None, create_betas=True, betas=None,
                 create_global_orient=True, global_orient=None, create_body_pose=True,
                 body_pose=None, create_transl=True, transl=None, dtype=torch.float32,
                 batch_size=1, joint_mapper=None, gender='neutral', vertex_ids=None, **kwargs):
        super().__init__()
        self.gender = gender
        if os.path.isdir(model_path):
            model_path = os.path.join(model_path, 'SMPL_{}.{ext}'.format(gender.upper(), ext='pkl'))
        with open(model_path, 'rb') as f:
            data = pickle.load(f, encoding='latin1')
        self.dtype = dtype
        self.batch_size = batch_size
        self.joint_mapper = joint_mapper
        self.vertex_joint_selector = VertexJointSelector(vertex_ids or VERTEX_IDS['smplh'])
        self.faces = np.asarray(data['f'])
        self.register_buffer('faces_tensor', torch.tensor(self.faces.astype(np.int64), dtype=torch.long))

        def param(name, create, value, width):
            if not create:
                return
            if value is None:
                value = torch.zeros([batch_size, width], dtype=dtype)
            elif not torch.is_tensor(value):
                value = torch.tensor(value, dtype=dtype)
            self.register_parameter(name, nn.Parameter(value, requires_grad=True))

        param('betas', create_betas, betas, self.NUM_BETAS)
        param('global_orient', create_global_orient, global_orient, 3)
        param('body_pose', create_body_pose, body_pose, self.NUM_BODY_JOINTS * 3)
        param('transl', create_transl, transl, 3)

        self.register_buffer('v_template', torch.tensor(_np(data['v_template']), dtype=dtype))
        self.register_buffer('shapedirs', torch.tensor(_np(data['shapedirs'])[:, :, :self.NUM_BETAS], dtype=dtype))
        self.register_buffer('J_regressor', torch.tensor(_np(data['J_regressor']), dtype=dtype))
        num_pose_basis = np.asarray(data['posedirs']).shape[-1]
        posedirs = np.reshape(_np(data['posedirs']), [-1, num_pose_basis]).T
        self.register_buffer('posedirs', torch.tensor(posedirs, dtype=dtype))
        parents = torch.tensor(_np(data['kintree_table'], np.int64)[0]).long()
        parents[0] = -1
        self.register_buffer('parents', parents)
        self.register_buffer('lbs_weights', torch.tensor(_np(data['weights']), dtype=dtype))

    def get_num_verts(self):
        return self.v_template.shape[0]

    def get_num_faces(self):
        return self.faces.shape[0]

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None,
                return_verts=True, return_full_pose=False, pose2rot=True, **kwargs):
        global_orient = global_orient if global_orient is not None else self.global_orient
        body_pose = body_pose if body_pose is not None else self.body_pose
        betas = betas if betas is not None else self.betas
        apply_trans = transl is not None or hasattr(self, 'transl')
        if transl is None and hasattr(self, 'transl'):
            transl = self.transl
        full_pose = torch.cat([global_orient, body_pose], dim=1)
        batch_size = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
        if betas.shape[0] != batch_size:
            betas = betas.expand(int(batch_size / betas.shape[0]), -1)
        vertices, joints = lbs(betas, full_pose, self.v_template, self.shapedirs, self.posedirs,
                               self.J_regressor, self.parents, self.lbs_weights,
                               pose2rot=pose2rot, dtype=self.dtype)
        joints = self.vertex_joint_selector(vertices, joints)
        if self.joint_mapper is not None:
            joints = self.joint_mapper(joints)
        if apply_trans:
            joints = joints + transl.unsqueeze(dim=1)
            vertices = vertices + transl.unsqueeze(dim=1)
        return ModelOutput(vertices=vertices if return_verts else None,
                           global_orient=global_orient, body_pose=body_pose, joints=joints,
                           betas=betas, full_pose=full_pose if return_full_pose else None)


def create(model_path, model_type='smpl', **kwargs):
    return SMPL(model_path, **kwargs)
