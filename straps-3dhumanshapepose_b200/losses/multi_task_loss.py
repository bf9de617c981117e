"""Homoscedastic-uncertainty multi-task loss (drop-in for reference losses/multi_task_loss.py:7-119), B200-native.

total = sum_t MSE_t * exp(-s_t) + s_t over t in {verts, joints2D (visible rows, labels mapped to [-1,1]), joints3D,
shape_params, pose_params}; the five log-variances s_t are nn.Parameters with the reference's names and initialisation, so
criterion.state_dict() checkpoints interchange.  Forward and the gradient seeds of all predictions are ONE fused reduction +
one elementwise kernel (csrc/loss.cu) instead of ~25 ATen launches.  CUDA tensors only (no CPU fallback).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

import config
from straps_b200 import _lib
from straps_b200._lib import StrapsError, check

_TASKS = ('verts', 'joints2D', 'joints3D', 'pose_params', 'shape_params')          # parameter registration order (reference)
_KERNEL_ORDER = ('verts', 'joints2D', 'joints3D', 'shape_params', 'pose_params')   # order of the C ABI
_OUTPUT_KEY = {'verts': 'verts', 'joints2D': 'joints2D', 'joints3D': 'joints3D', 'shape_params': 'shape_params',
               'pose_params': 'pose_params_rot_matrices'}


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_vars, vis, use_vis, sum_reduction, want_grads, *tensors):
        preds, targets = tensors[:5], tensors[5:]
        dev = log_vars.device
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        P = [p.contiguous() if p is not None else None for p in preds]
        T = [t.contiguous().float() if t is not None else None for t in targets]
        G = [torch.empty_like(p) if (p is not None and want_grads) else None for p in P]
        arr = lambda ts: (ctypes.c_void_p * 5)(*[t.data_ptr() if t is not None else None for t in ts])
        counts = (ctypes.c_int64 * 5)(*[p.numel() if p is not None else 0 for p in P])
        scratch = torch.empty(32, dtype=torch.float32, device=dev)
        out = torch.empty(6, dtype=torch.float32, device=dev)
        dlv = torch.empty(5, dtype=torch.float32, device=dev)
        lv = log_vars.contiguous()
        with torch.cuda.device(dev):
            check(_lib.lib().straps_multitask_loss(arr(P), arr(T), arr(G), counts, ptr(vis), use_vis, float(config.REGRESSOR_IMG_WH), ptr(lv),
                                                   1 if sum_reduction else 0, None, ptr(scratch), ptr(out), ptr(dlv),
                                                   ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), 'straps_multitask_loss')
        ctx.save_for_backward(dlv, *[g for g in G if g is not None])
        ctx.has_grad = [g is not None for g in G]
        total, parts = out[0].clone(), out[1:].clone()
        ctx.mark_non_differentiable(parts)
        return total, parts

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        dlv, *gs = ctx.saved_tensors
        it = iter(gs)
        grads = [next(it) * g_total if has else None for has in ctx.has_grad]
        return (dlv * g_total, None, None, None, None) + tuple(grads) + (None,) * 5


class HomoscedasticUncertaintyWeightedMultiTaskLoss(nn.Module):
    def __init__(self, losses_on, init_loss_weights=None, reduction='mean', eps=1e-6):
        super(HomoscedasticUncertaintyWeightedMultiTaskLoss, self).__init__()
        assert reduction in ['mean', 'sum'], "Invalid reduction for loss."
        self.losses_on = losses_on
        self.reduction = reduction
        for task in _TASKS:
            init = 0 if init_loss_weights is None else -np.log(init_loss_weights[task] + eps)
            setattr(self, task + '_log_var', nn.Parameter(torch.tensor(init).float(), requires_grad=task in losses_on))

    def forward(self, labels, outputs):
        on = [t for t in _KERNEL_ORDER if t in self.losses_on]
        if not on:
            return 0., {}
        preds = [outputs[_OUTPUT_KEY[t]] if t in on else None for t in _KERNEL_ORDER]
        targets = [labels[_OUTPUT_KEY[t]] if t in on else None for t in _KERNEL_ORDER]
        for t in preds + targets:
            if t is not None and not t.is_cuda:
                raise StrapsError('the multi-task loss runs on CUDA tensors only: the B200 path has no CPU fallback')
        vis = labels['vis'].contiguous() if 'vis' in labels.keys() else None
        if vis is not None and vis.dtype != torch.bool:
            vis = vis != 0
        log_vars = torch.stack([getattr(self, t + '_log_var') for t in _KERNEL_ORDER])
        want = torch.is_grad_enabled() and (log_vars.requires_grad or any(p is not None and p.requires_grad for p in preds))
        total, parts = _FusedLoss.apply(log_vars, vis, 1 if vis is not None else 0, self.reduction == 'sum', want, *preds, *targets)
        loss_dict = {t: parts[i] for i, t in enumerate(_KERNEL_ORDER) if t in on}
        # same key order as the reference builds its dict in
        loss_dict = {t: loss_dict[t] for t in ('verts', 'joints2D', 'joints3D', 'shape_params', 'pose_params') if t in loss_dict}
        return total, loss_dict
