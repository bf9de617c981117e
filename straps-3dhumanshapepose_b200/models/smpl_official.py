"""SMPL body model with the 90-joint superset, B200-native.

Drop-in for the reference's models/smpl_official.py:10-41 *and* for the third-party smplx.SMPL it subclasses
(not vendored by the reference; behaviour per SURVEY.md 8a S1-S8 / Appendix A): same constructor
(`SMPL(model_path, batch_size=...)`), same buffers/parameters, same forward keywords
(`betas, body_pose, global_orient, transl, pose2rot, return_verts, return_full_pose`) and the same
`ModelOutput` namedtuple.  The arithmetic -- Rodrigues / rotation-matrix input, shape + pose-corrective
blend shapes, kinematic chain, linear blend skinning over 6890 vertices, 24 + 21 + 45 joints -- is one fused
CUDA kernel plus a small joint-regression kernel (csrc/smpl.cu).
"""
import os
import pickle
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

import config
from straps_b200 import ops
from straps_b200.autograd import BatchRodrigues, SmplForward
from straps_b200._lib import StrapsError

ModelOutput = namedtuple('ModelOutput', ['vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose',
                                         'expression', 'left_hand_pose', 'right_hand_pose', 'jaw_pose'])
ModelOutput.__new__.__defaults__ = (None,) * len(ModelOutput._fields)

# smplx VertexJointSelector order: face (nose, reye, leye, rear, lear), feet (L big/small toe, heel, R ...),
# finger tips (left thumb..pinky, right thumb..pinky)  -- SURVEY.md 8a S7
EXTRA_JOINT_VERTEX_IDS = [332, 6260, 2800, 4071, 583,
                          3216, 3226, 3387, 6617, 6624, 6787,
                          2746, 2319, 2445, 2556, 2673,
                          6191, 5782, 5905, 6016, 6133]


class _ChumpyStub(object):
    """Lets an un-cleaned SMPL pickle (chumpy arrays) load without the chumpy package."""
    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {'x': state})

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self.__dict__.get('x', self.__dict__.get('a')))
        return a.astype(dtype) if dtype is not None else a


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split('.')[0] == 'chumpy':
            return _ChumpyStub
        return super(_Unpickler, self).find_class(module, name)


def _dense(a, dtype=np.float32):
    if hasattr(a, 'todense'):
        a = a.todense()
    return np.array(a, dtype=dtype)


class SMPL(nn.Module):
    NUM_JOINTS = 23
    NUM_BODY_JOINTS = 23
    NUM_BETAS = 10

    def __init__(self, model_path, batch_size=1, gender='neutral', dtype=torch.float32,
                 create_betas=True, betas=None, create_global_orient=True, global_orient=None,
                 create_body_pose=True, body_pose=None, create_transl=True, transl=None, **kwargs):
        super(SMPL, self).__init__()
        if dtype != torch.float32:
            raise NotImplementedError('the B200 SMPL kernels are fp32')
        if os.path.isdir(model_path):
            model_path = os.path.join(model_path, 'SMPL_%s.pkl' % gender.upper())
        with open(model_path, 'rb') as f:
            data = _Unpickler(f, encoding='latin1').load()
        self.gender, self.dtype, self.batch_size = gender, dtype, batch_size
        self.faces = np.asarray(data['f'])
        self.register_buffer('faces_tensor', torch.tensor(self.faces.astype(np.int64), dtype=torch.long))

        def default(name, create, value, width):
            if not create:
                return
            if value is None:
                value = torch.zeros([batch_size, width], dtype=dtype)
            elif not torch.is_tensor(value):
                value = torch.tensor(value, dtype=dtype)
            self.register_parameter(name, nn.Parameter(value, requires_grad=True))
        default('betas', create_betas, betas, self.NUM_BETAS)
        default('global_orient', create_global_orient, global_orient, 3)
        default('body_pose', create_body_pose, body_pose, self.NUM_BODY_JOINTS * 3)
        default('transl', create_transl, transl, 3)

        shapedirs = _dense(data['shapedirs'])[:, :, :self.NUM_BETAS]
        posedirs = _dense(data['posedirs'])
        posedirs = np.reshape(posedirs, [-1, posedirs.shape[-1]]).T            # [207, 20670], (vertex, xyz) inner
        parents = _dense(data['kintree_table'], np.int64)[0].copy()
        parents[0] = -1
        self.register_buffer('v_template', torch.tensor(_dense(data['v_template'])))
        self.register_buffer('shapedirs', torch.tensor(shapedirs))
        self.register_buffer('posedirs', torch.tensor(np.ascontiguousarray(posedirs)))
        self.register_buffer('J_regressor', torch.tensor(_dense(data['J_regressor'])))
        self.register_buffer('lbs_weights', torch.tensor(_dense(data['weights'])))
        self.register_buffer('parents', torch.tensor(parents, dtype=torch.long))
        # reference models/smpl_official.py:17-25
        for name, path in (('J_regressor_extra', config.J_REGRESSOR_EXTRA_PATH),
                           ('J_regressor_cocoplus', config.COCOPLUS_REGRESSOR_PATH),
                           ('J_regressor_h36m', config.H36M_REGRESSOR_PATH)):
            self.register_buffer(name, torch.tensor(np.load(path), dtype=torch.float32))
        self.register_buffer('extra_joints_idxs', torch.tensor(EXTRA_JOINT_VERTEX_IDS, dtype=torch.long))
        self._handles = {}

    def get_num_verts(self):
        return self.v_template.shape[0]

    def get_num_faces(self):
        return self.faces.shape[0]

    def _handle(self, device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key not in self._handles:
            cpu = lambda t: t.detach().cpu().numpy()
            extra = np.concatenate([cpu(self.J_regressor_extra), cpu(self.J_regressor_cocoplus), cpu(self.J_regressor_h36m)], 0)
            self._handles[key] = ops.SmplHandle(device, cpu(self.v_template), cpu(self.shapedirs), cpu(self.posedirs),
                                                cpu(self.J_regressor), cpu(self.lbs_weights), cpu(self.parents), extra,
                                                cpu(self.extra_joints_idxs))
        return self._handles[key]

    def _forward_with_grad(self, betas, body_pose, global_orient, transl, return_verts, return_full_pose, pose2rot):
        """Autograd path: the library's forward saves what csrc/smpl_bwd.cu needs; gradients reach every input that requires
        them -- rotation matrices (train/train_synthetic_otf_rendering.py:196-199), axis-angle poses through the Rodrigues
        backward kernel (pose2rot=True: `smpl_model(betas=pred_shape)`, train/...:206, runs on the module's own axis-angle
        parameters), betas and transl -- as smplx's graph would."""
        B = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
        if pose2rot:
            full = torch.cat([global_orient.reshape(-1, 1, 3), body_pose.reshape(-1, 23, 3)], dim=1)      # [B,24,3]
            if full.requires_grad:
                rotmats = BatchRodrigues.apply(full.reshape(-1, 3)).view(-1, 24, 3, 3)
            else:
                rotmats = ops.batch_rodrigues(full.reshape(-1, 3)).view(-1, 24, 3, 3)
        else:
            rotmats = torch.cat([global_orient.reshape(-1, 1, 3, 3), body_pose.reshape(-1, 23, 3, 3)], dim=1)
        if rotmats.shape[0] != B:
            rotmats = rotmats.expand(B, -1, -1, -1)
        if betas.shape[0] != B:
            betas = betas.expand(B, -1)
        verts, joints = SmplForward.apply(self._handle(betas.device), rotmats, betas)
        if transl is not None:
            if transl.shape[0] not in (1, B):
                raise StrapsError('SMPL: transl has batch %d but the call has batch %d -- construct SMPL(batch_size=%d) '
                                  '(reference run_train.py:109-110)' % (transl.shape[0], B, B))
            verts = verts + transl.unsqueeze(1)
            joints = joints + transl.unsqueeze(1)
        full_pose = torch.cat([global_orient, body_pose], dim=1) if return_full_pose else None
        return ModelOutput(vertices=verts if return_verts else None, global_orient=global_orient, body_pose=body_pose,
                           joints=joints, betas=betas, full_pose=full_pose)

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, return_verts=True,
                return_full_pose=False, pose2rot=True, **kwargs):
        # `get_skin` and friends arrive through **kwargs and are ignored, as in smplx (smpl_official.py:28)
        global_orient = global_orient if global_orient is not None else self.global_orient
        body_pose = body_pose if body_pose is not None else self.body_pose
        betas = betas if betas is not None else self.betas
        if transl is None and hasattr(self, 'transl'):
            transl = self.transl
        if not betas.is_cuda:
            raise StrapsError('SMPL.forward needs CUDA tensors: the B200 path has no CPU fallback')
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (betas, body_pose, global_orient, transl)):
            return self._forward_with_grad(betas, body_pose, global_orient, transl, return_verts, return_full_pose, pose2rot)
        verts, joints = self._handle(betas.device).forward(global_orient.detach(), body_pose.detach(), betas.detach(),
                                                           transl.detach() if transl is not None else None, pose2rot)
        full_pose = torch.cat([global_orient, body_pose], dim=1) if return_full_pose else None
        return ModelOutput(vertices=verts if return_verts else None, global_orient=global_orient, body_pose=body_pose,
                           joints=joints, betas=betas, full_pose=full_pose)
