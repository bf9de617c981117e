"""Homoscedastic-uncertainty multi-task loss (drop-in for reference losses/multi_task_loss.py:7-119).

total = sum_t MSE_t * exp(-s_t) + s_t over t in {verts, joints2D (visible rows, labels mapped to [-1,1]),
joints3D, shape_params, pose_params}; the five log-variances s_t are nn.Parameters with the reference's
names and initialisation, so criterion.state_dict() checkpoints interchange.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import config

_TASKS = ('verts', 'joints2D', 'joints3D', 'pose_params', 'shape_params')


class HomoscedasticUncertaintyWeightedMultiTaskLoss(nn.Module):
    def __init__(self, losses_on, init_loss_weights=None, reduction='mean', eps=1e-6):
        super(HomoscedasticUncertaintyWeightedMultiTaskLoss, self).__init__()
        assert reduction in ['mean', 'sum'], "Invalid reduction for loss."
        self.losses_on = losses_on
        self.reduction = reduction
        for task in _TASKS:
            init = 0 if init_loss_weights is None else -np.log(init_loss_weights[task] + eps)
            setattr(self, task + '_log_var', nn.Parameter(torch.tensor(init).float(), requires_grad=task in losses_on))

    def _term(self, task, pred, target):
        raw = F.mse_loss(pred, target, reduction=self.reduction)
        log_var = getattr(self, task + '_log_var')
        weighted = raw * torch.exp(-log_var)
        return weighted + log_var, weighted

    def forward(self, labels, outputs):
        total_loss, loss_dict = 0., {}
        if 'verts' in self.losses_on:
            t, loss_dict['verts'] = self._term('verts', outputs['verts'], labels['verts'])
            total_loss += t
        if 'joints2D' in self.losses_on:
            label, pred = labels['joints2D'], outputs['joints2D']
            if 'vis' in labels.keys():
                label, pred = label[labels['vis'], :], pred[labels['vis'], :]
            label = (2.0 * label) / config.REGRESSOR_IMG_WH - 1.0
            t, loss_dict['joints2D'] = self._term('joints2D', pred, label)
            total_loss += t
        if 'joints3D' in self.losses_on:
            t, loss_dict['joints3D'] = self._term('joints3D', outputs['joints3D'], labels['joints3D'])
            total_loss += t
        if 'shape_params' in self.losses_on:
            t, loss_dict['shape_params'] = self._term('shape_params', outputs['shape_params'], labels['shape_params'])
            total_loss += t
        if 'pose_params' in self.losses_on:
            t, loss_dict['pose_params'] = self._term('pose_params', outputs['pose_params_rot_matrices'],
                                                     labels['pose_params_rot_matrices'])
            total_loss += t
        return total_loss, loss_dict
