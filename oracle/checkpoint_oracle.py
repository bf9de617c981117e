"""ORACLE / TEST INFRASTRUCTURE -- a reference-style training checkpoint, rebuilt without the reference.

SURVEY.md 8f row N3: the reference writes `.tar` checkpoints with
    {'epoch', 'best_epoch', 'best_epoch_val_metrics', 'model_state_dict', 'best_model_state_dict',
     'optimiser_state_dict', 'criterion_state_dict'}                       (train/train_synthetic_otf_rendering.py:365-377)
where the optimiser is `optim.Adam(list(regressor.parameters()) + list(criterion.parameters()), lr)` (run_train.py:200-201), i.e.
the optimiser state is numbered over ALL 71 tensors, the log-variances of switched-off losses included (requires_grad=False,
losses/multi_task_loss.py:47-71: they never get a state entry).  A real checkpoint is ~190 MB, far too big for a fixture, so:

* oracle/gen_checkpoint_fixture.py runs the UNMODIFIED reference modules (regressor, SMPL on the smplx shim, loss, torch.optim.Adam)
  for a few steps, saves the checkpoint's METADATA (keys, shapes, dtypes, float64 checksums of every tensor, param_groups, state
  indices, step counts) to tests/golden/checkpoint_ref.json, and asserts that `build()` below reproduces every tensor of the real
  checkpoint (same seeds, the functional oracle pinned bit-identical to the reference);
* the tests call `build()` (here, or on the GPU box where /root/reference does not exist), check it against the committed metadata,
  write it with torch.save and resume the B200 regressor + DataParallelAdam + criterion from it.
"""
import os

import numpy as np
import torch

import straps_oracle as O

C_IN, BATCH, STEPS, LR = 17, 2, 2, 1e-4
LOSSES_ON = ['verts', 'joints2D', 'shape_params']            # a subset: joints3D / pose_params log-variances stay frozen
LOSS_WEIGHTS = {'verts': 1.0, 'joints2D': 0.1, 'pose_params': 0.1, 'shape_params': 0.1, 'joints3D': 1.0}   # run_train.py:53-54
CRITERION_ORDER = ['verts', 'joints2D', 'joints3D', 'pose_params', 'shape_params']   # registration order, losses/multi_task_loss.py:47-71
WEIGHT_SEED = 12


def step_data(step, smpl_oracle):
    """Seeded inputs and targets of training step `step` (targets from the oracle SMPL, as train/...:121-145 makes them)."""
    from straps_b200 import synthetic_inputs
    rng = np.random.RandomState(900 + step)
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(BATCH, C_IN, seed=40 + step))
    betas = torch.from_numpy(rng.normal(0, 1, (BATCH, 10)).astype(np.float32))
    with torch.no_grad():
        R = O.rot6d_to_rotmat(torch.from_numpy(rng.normal(0, 1, (BATCH, 144)).astype(np.float32))).view(BATCH, 24, 3, 3)
        v, j = smpl_oracle.forward_rotmats(R, betas)
    j2d = torch.from_numpy(rng.uniform(-20, 276, (BATCH, 17, 2)).astype(np.float32))
    return x, {'verts': v, 'joints2D': j2d, 'joints3D': j[:, O.ALL_JOINTS_TO_H36M_MAP][:, O.H36M_TO_J14], 'shape_params': betas,
               'pose_params_rot_matrices': R}


def regressor_param_names(sd):
    """nn.Module.parameters() order of the reference regressor: registration order with the duplicated IEF keys removed
    (ief_layers.* alias fc1/2/3: models/ief_module.py:16-28) and the BatchNorm buffers skipped."""
    # the modules register conv1, bn1, then the blocks (conv1, bn1, conv2, bn2, downsample.0, downsample.1): rebuild that order
    out = ['image_encoder.conv1.weight', 'image_encoder.bn1.weight', 'image_encoder.bn1.bias']
    for name, cin, cout, stride, ds in O._BLOCKS:
        p = 'image_encoder.' + name
        out += [p + '.conv1.weight', p + '.bn1.weight', p + '.bn1.bias', p + '.conv2.weight', p + '.bn2.weight', p + '.bn2.bias']
        if ds:
            out += [p + '.downsample.0.weight', p + '.downsample.1.weight', p + '.downsample.1.bias']
    out += ['ief_module.fc1.weight', 'ief_module.fc1.bias', 'ief_module.fc2.weight', 'ief_module.fc2.bias',
            'ief_module.fc3.weight', 'ief_module.fc3.bias']
    assert all(n in sd for n in out) and len(out) == 66
    return out


def build(additional_dir, steps=STEPS):
    """-> (checkpoint dict in the reference's format after `steps` training steps, the 71 optimiser parameter names in order)."""
    smpl = O.SmplOracle(additional_dir, batch_size=BATCH)
    init = O.load_initial_params(os.path.join(additional_dir, 'neutral_smpl_mean_params_6dpose.npz'))
    sd = O.make_regressor_state(C_IN, seed=WEIGHT_SEED)
    names = regressor_param_names(sd)
    live = {k: v.clone() for k, v in sd.items() if 'ief_layers' not in k}
    for n in names:
        live[n].requires_grad_(True)
    for i, n in ((0, 'fc1'), (2, 'fc2'), (4, 'fc3')):
        live['ief_module.ief_layers.%d.weight' % i] = live['ief_module.%s.weight' % n]
        live['ief_module.ief_layers.%d.bias' % i] = live['ief_module.%s.bias' % n]
    lv = {t: v.clone().requires_grad_(t in LOSSES_ON) for t, v in O.init_log_vars(LOSS_WEIGHTS).items()}
    opt = torch.optim.Adam([live[n] for n in names] + [lv[t] for t in CRITERION_ORDER], lr=LR)
    for step in range(steps):
        x, tg = step_data(step, smpl)
        tg['vis'] = O.joints2d_visibility(tg['joints2D'])
        stats = {}
        opt.zero_grad()
        o = O.regress_and_pose(x, live, init, smpl, train=True, stats_out=stats)
        outs = {'verts': o['vertices'], 'joints2D': o['joints2d_coco'], 'joints3D': o['joints_h36mlsp'], 'shape_params': o['shape'],
                'pose_params_rot_matrices': o['rotmats']}
        loss, _ = O.multi_task_loss(tg, outs, lv, losses_on=LOSSES_ON)
        loss.backward()
        opt.step()
        with torch.no_grad():
            for k, v in stats.items():                           # running statistics of this step (momentum 0.1)
                live['image_encoder.' + k].copy_(v)
            for k in live:
                if k.endswith('num_batches_tracked'):
                    live[k] += 1
    model_sd = {k: live[k].detach().clone() for k in sd}         # the reference's 132 keys, in its order
    ckpt = {'epoch': 3, 'best_epoch': 2, 'best_epoch_val_metrics': {'pves': 0.123, 'mpjpes_pa': 0.045},
            'model_state_dict': model_sd, 'best_model_state_dict': {k: v.clone() for k, v in model_sd.items()},
            'optimiser_state_dict': opt.state_dict(),
            'criterion_state_dict': {t + '_log_var': lv[t].detach().clone() for t in CRITERION_ORDER}}
    return ckpt, names + [t + '_log_var' for t in CRITERION_ORDER]


def metadata(ckpt):
    """JSON-able description of a checkpoint dict: structure + float64 checksums of every tensor."""
    def desc(t):
        a = t.detach().cpu().double()
        return {'shape': list(t.shape), 'dtype': str(t.dtype).replace('torch.', ''), 'sum': float(a.sum()), 'abs_sum': float(a.abs().sum())}
    osd = ckpt['optimiser_state_dict']
    return {'keys': list(ckpt.keys()), 'epoch': ckpt['epoch'], 'best_epoch': ckpt['best_epoch'],
            'best_epoch_val_metrics': ckpt['best_epoch_val_metrics'],
            'model_state_dict': {k: desc(v) for k, v in ckpt['model_state_dict'].items()},
            'criterion_state_dict': {k: desc(v) for k, v in ckpt['criterion_state_dict'].items()},
            'optimiser_param_groups': [{k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in g.items()}
                                       for g in osd['param_groups']],
            'optimiser_state': {str(i): {'step': float(st['step']), 'exp_avg': desc(st['exp_avg']), 'exp_avg_sq': desc(st['exp_avg_sq'])}
                                for i, st in osd['state'].items()}}
