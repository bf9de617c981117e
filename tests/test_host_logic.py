"""CPU: host-side logic of the drop-in modules (no kernels run)."""
import os

import numpy as np
import pytest
import torch

import straps_oracle as O
from conftest import SCRATCH
from straps_b200 import synthetic_assets, synthetic_inputs
from straps_b200._lib import StrapsError


def test_state_dict_layout_is_the_reference_one(assets_root):
    from models.regressor import SingleInputRegressor
    reg = SingleInputRegressor(18, 18, 3)
    sd = O.make_regressor_state(18, seed=1)
    assert list(reg.state_dict().keys()) == list(sd.keys()) or set(reg.state_dict().keys()) == set(sd.keys())
    assert len(reg.state_dict()) == 132
    reg.load_state_dict(sd)
    assert reg.state_dict()['image_encoder.conv1.weight'].shape == (64, 18, 7, 7)
    assert reg.ief_module.fc1.weight.data_ptr() == reg.ief_module.ief_layers[0].weight.data_ptr()
    assert reg.ief_module.initial_params_estimate.shape == (157,)
    assert float(reg.ief_module.initial_params_estimate[0]) == pytest.approx(0.9)
    n = sum(p.numel() for p in reg.image_encoder.parameters())
    assert n == 11223552          # SURVEY.md section 0.5 (C=18)
    with pytest.raises(NotImplementedError):
        SingleInputRegressor(18, 50, 3)


def test_no_cpu_fallback(assets_root):
    import config
    from models.regressor import SingleInputRegressor
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat
    reg = SingleInputRegressor(17, 18, 3).eval()
    with pytest.raises(StrapsError):
        reg(torch.zeros(1, 17, 256, 256))
    with pytest.raises(StrapsError):
        rot6d_to_rotmat(torch.zeros(2, 144))
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=2)
    with pytest.raises(StrapsError):
        with torch.no_grad():
            smpl(betas=torch.zeros(2, 10))
    assert [n for n, _ in smpl.named_parameters()] == ['betas', 'global_orient', 'body_pose', 'transl']
    assert smpl.faces_tensor.dtype == torch.int64 and smpl.parents.dtype == torch.int64
    assert tuple(smpl.posedirs.shape) == (207, 20670)


def test_dataset_matches_reference_semantics():
    from data.synthetic_training_dataset import SyntheticTrainingDataset
    path = synthetic_assets.write_synthetic_dataset(os.path.join(SCRATCH, 'ds.npz'), n=16, seed=0)
    ds = SyntheticTrainingDataset(path)
    assert len(ds) == 16
    item = ds[torch.tensor(3)]
    assert item['pose'].shape == (72,) and item['shape'].shape == (10,) and item['pose'].dtype == torch.float32
    assert len(SyntheticTrainingDataset(path, 'h36m')) == 4
    assert len(SyntheticTrainingDataset(path, 'not_amass')) == 12
    with pytest.raises(AssertionError):
        SyntheticTrainingDataset(path, 'coco')


def test_visibility_mask_and_proxy_inputs():
    from utils.joints2d_utils import check_joints2d_visibility_torch
    j = torch.tensor([[[0., 0.], [256., 256.], [256.5, 3.], [-0.1, 9.], [9., -1.], [128., 300.]]])
    assert check_joints2d_visibility_torch(j, 256).tolist() == [[True, True, False, False, False, False]]
    assert torch.equal(check_joints2d_visibility_torch(j, 256), O.joints2d_visibility(j))
    x = synthetic_inputs.make_proxy_batch(2, 17, seed=3)
    assert x.shape == (2, 17, 256, 256) and x.dtype == np.float32
    assert set(np.unique(x[:, 0])) <= {0.0, 1.0}
    assert all(int((x[0, c] != 0).sum()) == 256 for c in range(1, 17))
    assert np.array_equal(x, synthetic_inputs.make_proxy_batch(2, 17, seed=3))


def test_conv1_pair_layout_scheme_matches_conv2d():
    """tools/conv1_s2d_emulation.py: conv1_s2d_kernel's pixel-pair layout, box / window / K-step arithmetic and the fused max-pool
    epilogue, replayed in numpy, equal the 7x7 / stride 2 / pad 3 conv2d followed by max_pool2d(3, 2, 1)."""
    import subprocess
    import sys
    from conftest import REPO
    res = subprocess.run([sys.executable, os.path.join(REPO, 'tools', 'conv1_s2d_emulation.py')], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and 'pair-layout conv1 == conv2d' in res.stdout, res.stdout[-1000:] + res.stderr[-1000:]


def test_select_joints_is_the_reference_list_indexing():
    """straps_b200.ops.select_joints (one cached index_select, graph-capturable) == the reference's chained list indexing of the joint
    superset (train/train_synthetic_otf_rendering.py:207-215), values and gradients."""
    import config
    from straps_b200.ops import select_joints
    rng = np.random.RandomState(0)
    j = torch.from_numpy(rng.normal(0, 1, (3, 90, 3)).astype(np.float32)).requires_grad_(True)
    j2 = j.detach().clone().requires_grad_(True)
    a = select_joints(j, config.ALL_JOINTS_TO_H36M_MAP, config.H36M_TO_J14)
    b = j2[:, config.ALL_JOINTS_TO_H36M_MAP, :][:, config.H36M_TO_J14, :]
    assert torch.equal(a, b)
    c = select_joints(j, config.ALL_JOINTS_TO_COCO_MAP)
    assert torch.equal(c, j2[:, config.ALL_JOINTS_TO_COCO_MAP, :])
    w = torch.from_numpy(rng.normal(0, 1, tuple(a.shape)).astype(np.float32))
    (a * w).sum().backward()
    (b * w).sum().backward()
    assert torch.allclose(j.grad, j2.grad, atol=1e-6)


def test_early_all_reduce_range_is_layer4_plus_ief(assets_root):
    """DataParallelAdam.early_range_of: in the reference's parameter order (regressor.parameters() + criterion.parameters()) layer4 and the
    IEF module are one contiguous run at the deep end of the flat bucket -- what the overlapped all-reduce relies on -- and any other order
    is refused."""
    from models.regressor import SingleInputRegressor
    from losses.multi_task_loss import HomoscedasticUncertaintyWeightedMultiTaskLoss as Loss
    from straps_b200.parallel import DataParallelAdam
    reg = SingleInputRegressor(17, 18, 3)
    crit = Loss(['verts', 'joints2D'], init_loss_weights=None)
    opt = DataParallelAdam(list(reg.parameters()) + list(crit.parameters()), broadcast=False)
    lo, hi = opt.early_range_of(reg)
    n_l4 = sum(p.numel() for p in reg.image_encoder.layer4.parameters())
    n_ief = sum(p.numel() for p in reg.ief_module.parameters())
    n_enc = sum(p.numel() for p in reg.image_encoder.parameters())
    assert hi - lo == n_l4 + n_ief == 9079965 and lo == n_enc - n_l4 and hi == n_enc + n_ief
    assert not opt.enable_overlap(reg)                      # one process, CPU bucket: nothing to overlap
    reg2 = SingleInputRegressor(17, 18, 3)
    shuffled = list(reg2.ief_module.parameters()) + list(reg2.image_encoder.parameters())
    assert DataParallelAdam(shuffled, broadcast=False).early_range_of(reg2) is None
