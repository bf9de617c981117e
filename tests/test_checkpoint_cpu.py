"""CPU: SURVEY.md 8f row N3 -- the reference's checkpoint dict round-trips through the flat-bucket optimiser and torch.optim.Adam."""
import os

import numpy as np
import pytest
import torch

from conftest import SCRATCH


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))


def _step(model, opt, x, y):
    opt.zero_grad()
    torch.nn.functional.mse_loss(model(x), y).backward()
    opt.step()


def test_optimiser_state_round_trips_with_torch_adam(assets_root):
    from straps_b200.parallel import DataParallelAdam
    from utils.checkpoint_utils import save_checkpoint, resume_from_checkpoint, load_training_info_from_checkpoint, CHECKPOINT_KEYS
    from losses.multi_task_loss import HomoscedasticUncertaintyWeightedMultiTaskLoss as Loss
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    crit = Loss(['verts', 'joints2D'], init_loss_weights={'verts': 1.0, 'joints2D': 0.1, 'joints3D': 1.0, 'pose_params': 0.1, 'shape_params': 0.1})

    # A: flat-bucket Adam for 2 steps -> checkpoint -> torch Adam continues;  B: torch Adam all the way
    ma, mb = _model(3), _model(3)
    oa = DataParallelAdam(ma.parameters(), lr=1e-2)
    ob = torch.optim.Adam(mb.parameters(), lr=1e-2)
    for _ in range(2):
        _step(ma, oa, x, y)
        _step(mb, ob, x, y)
    path = os.path.join(SCRATCH, 'ckpt_epoch3.tar')
    os.makedirs(SCRATCH, exist_ok=True)
    save_checkpoint(path, 3, 2, {'pves': 0.1, 'old_metric': 0.5}, ma, ma.state_dict(), oa, crit)
    ck = torch.load(path, weights_only=False)
    assert tuple(ck.keys()) == CHECKPOINT_KEYS
    assert set(ck['optimiser_state_dict']['state'].keys()) == {0, 1, 2, 3}
    assert set(ob.state_dict()['param_groups'][0].keys()) <= set(ck['optimiser_state_dict']['param_groups'][0].keys()) | {'params'}
    mc = _model(99)
    oc = torch.optim.Adam(mc.parameters(), lr=5.0)         # wrong lr on purpose: the checkpoint must restore it
    resume_from_checkpoint(path, mc, oc, crit)
    _step(mc, oc, x, y)
    _step(mb, ob, x, y)
    for pc, pb in zip(mc.parameters(), mb.parameters()):
        assert torch.allclose(pc, pb, rtol=1e-5, atol=1e-6)

    # and the other way round: a torch.optim.Adam checkpoint resumes in the flat-bucket optimiser
    torch.save({'epoch': 0, 'best_epoch': 0, 'best_epoch_val_metrics': {}, 'model_state_dict': mb.state_dict(),
                'best_model_state_dict': mb.state_dict(), 'optimiser_state_dict': ob.state_dict(), 'criterion_state_dict': crit.state_dict()}, path)
    md = _model(7)
    od = DataParallelAdam(md.parameters(), lr=5.0)
    resume_from_checkpoint(path, md, od, crit)
    od.bucket.bump_versions()
    assert od.step_count == 3 and od.lr == pytest.approx(1e-2)
    _step(md, od, x, y)
    _step(mb, ob, x, y)
    for pd, pb in zip(md.parameters(), mb.parameters()):
        assert torch.allclose(pd, pb, rtol=1e-5, atol=1e-6)

    cur, best, wts, metrics = load_training_info_from_checkpoint({'epoch': 3, 'best_epoch': 2, 'best_model_state_dict': {},
                                                                 'best_epoch_val_metrics': {'pves': 0.1, 'old_metric': 0.5}}, ['pves', 'mpjpes_pa'])
    assert cur == 4 and best == 2 and metrics == {'pves': 0.1, 'mpjpes_pa': np.inf}


def test_load_rejects_mismatched_state(assets_root):
    from straps_b200.parallel import DataParallelAdam
    m = _model(1)
    o = DataParallelAdam(m.parameters())
    sd = torch.optim.Adam(_model(1)[0:1].parameters()).state_dict()
    with pytest.raises(ValueError):
        o.load_state_dict(sd)


def _check_against_fixture(ckpt, names):
    """The rebuilt checkpoint equals the one the UNMODIFIED reference wrote (tests/golden/checkpoint_ref.json, metadata only)."""
    import json
    import checkpoint_oracle as CK
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, 'checkpoint_ref.json')))
    got = json.loads(json.dumps(CK.metadata(ckpt)))
    assert names == ref['param_names'] and got['keys'] == ref['keys']
    assert got['optimiser_param_groups'] == ref['optimiser_param_groups']
    assert sorted(got['optimiser_state']) == sorted(ref['optimiser_state'])

    def close(a, b, what, tol=1e-4):
        assert a['shape'] == b['shape'] and a['dtype'] == b['dtype'], what
        assert abs(a['abs_sum'] - b['abs_sum']) <= tol * max(abs(b['abs_sum']), 1e-12), (what, a, b)     # another CPU may round differently
    for k, v in ref['model_state_dict'].items():
        close(got['model_state_dict'][k], v, k)
    for k, v in ref['criterion_state_dict'].items():
        close(got['criterion_state_dict'][k], v, k)
    for i, st in ref['optimiser_state'].items():
        assert got['optimiser_state'][i]['step'] == st['step']
        # Adam moments come from ONE backward pass of the CPU reference: conv1's weight gradient is a heavily cancelling sum whose
        # rounding depends on the host's vector width and thread count (2.6e-4 seen between this container and the GPU box's Xeon)
        close(got['optimiser_state'][i]['exp_avg'], st['exp_avg'], 'exp_avg %s' % i, tol=2e-3)
        close(got['optimiser_state'][i]['exp_avg_sq'], st['exp_avg_sq'], 'exp_avg_sq %s' % i, tol=2e-3)
    return ref


def test_reference_checkpoint_with_frozen_log_variances_resumes(assets_root, additional_dir):
    """SURVEY.md 8f row N3 for real: a checkpoint in the reference's format -- SingleInputRegressor + optim.Adam(regressor U criterion)
    + criterion with a SUBSET of losses on (run_train.py:194-209, losses/multi_task_loss.py:47-71), reproduced bit for bit from the
    reference's own run by oracle/checkpoint_oracle.py -- loads into the drop-in regressor, criterion and DataParallelAdam, and what
    DataParallelAdam writes back loads into torch.optim.Adam.  (The GPU half, tests/test_gpu_checkpoint.py, also takes the next step.)"""
    import checkpoint_oracle as CK
    from models.regressor import SingleInputRegressor
    from losses.multi_task_loss import HomoscedasticUncertaintyWeightedMultiTaskLoss as Loss
    from straps_b200.parallel import DataParallelAdam
    from utils.checkpoint_utils import resume_from_checkpoint, save_checkpoint
    ckpt, names = CK.build(additional_dir)
    ref = _check_against_fixture(ckpt, names)
    path = os.path.join(SCRATCH, 'reference_style_epoch3.tar')
    os.makedirs(SCRATCH, exist_ok=True)
    torch.save(ckpt, path)

    reg = SingleInputRegressor(CK.C_IN, 18, 3)
    crit = Loss(CK.LOSSES_ON, init_loss_weights=None)           # wrong initial log-variances on purpose: the checkpoint restores them
    params = list(reg.parameters()) + list(crit.parameters())
    assert [n for n, _ in reg.named_parameters()] + [n for n, _ in crit.named_parameters()] == names
    opt = DataParallelAdam(params, lr=5.0)
    assert len(opt.bucket.all_params) == 71 and len(opt.bucket.plist) == 69          # the two frozen log-variances keep their index
    resume_from_checkpoint(path, reg, opt, crit)
    assert opt.step_count == CK.STEPS and opt.lr == pytest.approx(CK.LR)
    osd = ckpt['optimiser_state_dict']
    for i, p, o in zip(opt.bucket.index, opt.bucket.plist, opt.bucket.offsets):
        assert torch.equal(opt.exp_avg[o:o + p.numel()].view(p.shape), osd['state'][i]['exp_avg']), names[i]
        assert torch.equal(opt.exp_avg_sq[o:o + p.numel()].view(p.shape), osd['state'][i]['exp_avg_sq']), names[i]
    for k, v in ckpt['model_state_dict'].items():
        assert torch.equal(reg.state_dict()[k], v), k
    for k, v in ckpt['criterion_state_dict'].items():
        assert torch.equal(crit.state_dict()[k], v), k
    frozen = [names.index('joints3D_log_var'), names.index('pose_params_log_var')]
    assert all(str(i) not in ref['optimiser_state'] for i in frozen)

    # and back: what the drop-in writes is what torch.optim.Adam (i.e. the reference) reads
    out = os.path.join(SCRATCH, 'b200_written_epoch3.tar')
    save_checkpoint(out, 3, 2, ckpt['best_epoch_val_metrics'], reg, reg.state_dict(), opt, crit)
    back = torch.load(out, weights_only=False)
    assert sorted(back['optimiser_state_dict']['state']) == sorted(osd['state'])
    assert back['optimiser_state_dict']['param_groups'][0]['params'] == list(range(71))
    t_opt = torch.optim.Adam([torch.nn.Parameter(p.detach().clone(), requires_grad=p.requires_grad) for p in params], lr=1.0)
    t_opt.load_state_dict(back['optimiser_state_dict'])
    assert t_opt.param_groups[0]['lr'] == pytest.approx(CK.LR)
    for i, st in osd['state'].items():
        assert torch.equal(t_opt.state[t_opt.param_groups[0]['params'][i]]['exp_avg'], st['exp_avg'])
    # a checkpoint trained with ANOTHER losses_on must be refused loudly, not silently misplaced
    crit_all = Loss(['verts', 'joints2D', 'joints3D', 'pose_params', 'shape_params'])
    reg2 = SingleInputRegressor(CK.C_IN, 18, 3)
    opt_all = DataParallelAdam(list(reg2.parameters()) + list(crit_all.parameters()))
    opt_all.load_state_dict(osd)                                  # superset of trainable parameters: fine, missing states stay zero
    crit_few = Loss(['verts'])
    opt_few = DataParallelAdam(list(reg2.parameters()) + list(crit_few.parameters()))
    with pytest.raises(ValueError):
        opt_few.load_state_dict(osd)
