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
