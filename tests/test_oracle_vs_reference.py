"""CPU, authoring container only: pin the oracle against the UNMODIFIED reference modules imported from
/root/reference (skipped on the GPU box, where the reference tree does not exist)."""
import os

import numpy as np
import pytest
import torch

import ref_harness
import straps_oracle as O
from conftest import rel_err
from straps_b200 import synthetic_inputs

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason='/root/reference not present')


@pytest.fixture(scope='module')
def ref(assets_root):
    return ref_harness.load_reference(assets_root)


@pytest.mark.parametrize('C,B', [(17, 3), (18, 1), (1, 2)])
def test_encoder_ief_smpl_bit_identical_to_reference(ref, C, B, smpl_oracle, additional_dir):
    sd = O.make_regressor_state(C, seed=5)
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=9))
    with ref.cwd(), torch.no_grad():
        reg = ref.SingleInputRegressor(C, 18, 3)
        assert set(reg.state_dict().keys()) == set(sd.keys()) and len(sd) == 132
        reg.load_state_dict(sd)
        reg.eval()
        cam, pose, shape = reg(x)
        R = ref.rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
        smpl = ref.SMPL(ref.config.SMPL_MODEL_DIR, batch_size=B)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
    init = O.load_initial_params(os.path.join(additional_dir, 'neutral_smpl_mean_params_6dpose.npz'))
    with torch.no_grad():
        o = O.regress_and_pose(x, sd, init, smpl_oracle)
    assert torch.equal(o['cam'], cam) and torch.equal(o['pose6d'], pose) and torch.equal(o['shape'], shape)
    assert torch.equal(o['rotmats'], R)
    assert torch.equal(o['vertices'], out.vertices) and torch.equal(o['joints'], out.joints)


def test_train_mode_batchnorm_matches_reference(ref):
    C, B = 17, 4
    sd = O.make_regressor_state(C, seed=6)
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=2))
    with ref.cwd(), torch.no_grad():
        reg = ref.SingleInputRegressor(C, 18, 3)
        reg.load_state_dict(sd)
        reg.train()
        feat = reg.image_encoder(x)
    stats = {}
    with torch.no_grad():
        f2 = O.encoder_forward(x, sd, train=True, stats_out=stats)
    assert rel_err(f2.numpy(), feat.numpy()) < 1e-6
    new = reg.state_dict()
    for k, v in stats.items():
        assert rel_err(v.numpy(), new['image_encoder.' + k].numpy()) < 1e-6


def test_product_constants_equal_reference(ref):
    import config as mine
    for k in ('SMPL_MODEL_DIR', 'FOCAL_LENGTH', 'REGRESSOR_IMG_WH', 'ALL_JOINTS_TO_COCO_MAP', 'ALL_JOINTS_TO_H36M_MAP',
              'H36M_TO_J17', 'H36M_TO_J14'):
        a, b = getattr(mine, k), getattr(ref.config, k)
        if k == 'SMPL_MODEL_DIR':
            assert a.endswith(b)
        else:
            assert a == b, k
    assert O.ALL_JOINTS_TO_COCO_MAP == ref.config.ALL_JOINTS_TO_COCO_MAP
    assert O.H36M_TO_J14 == ref.config.H36M_TO_J14


def test_heatmap_input_generator_matches_reference(ref):
    """The synthetic proxy inputs paste exactly the reference's 16x16 truncated Gaussian."""
    j = torch.tensor([[[100., 60.], [30., 200.]]])
    hm = ref.heatmaps(j, 256)[0].numpy()
    g = synthetic_inputs._gaussian16()
    assert np.allclose(hm[0, 60 - 8:60 + 8, 100 - 8:100 + 8], g, atol=1e-7)
    assert abs(g.max() - 0.982) < 1e-3
