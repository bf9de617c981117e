"""GPU parity: ResNet-18 encoder + IEF (through models.regressor.SingleInputRegressor -> ctypes -> C ABI)
against the CPU oracle and the reference-generated fixtures, for both convolution modes."""
import os

import numpy as np
import pytest
import torch

import straps_oracle as O
from conftest import golden, rel_err, rel_l2, RTOL, WEIGHT_SEED, INPUT_SEED
from straps_b200 import synthetic_inputs

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
MODES = ['fp32_simt', 'f16x3_tc']


def _regressor(C, mode, sd):
    from models.regressor import SingleInputRegressor
    reg = SingleInputRegressor(C, 18, 3, conv_mode=mode)
    reg.load_state_dict(sd)
    return reg.to(DEV).eval()


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('C', [17, 18])
def test_config2_b2_against_reference_fixture(mode, C, assets_root):
    """encoder + 3x IEF + rot6d + SMPL + projection, B=2, vs outputs of the UNMODIFIED reference."""
    import config
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat
    from utils.cam_utils import orthographic_project_torch
    g = golden('regressor_c%d_b2.npz' % C)
    sd = O.make_regressor_state(C, seed=WEIGHT_SEED)
    reg = _regressor(C, mode, sd)
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=2).to(DEV)
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(2, C, seed=INPUT_SEED)).to(DEV)
    with torch.no_grad():
        feat = reg.image_encoder(x)
        cam, pose, shape = reg(x)
        assert not pose.is_contiguous() and cam.shape == (2, 3) and pose.shape == (2, 144) and shape.shape == (2, 10)
        R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
        j2d = orthographic_project_torch(out.joints[:, config.ALL_JOINTS_TO_COCO_MAP, :], cam)
        j_lsp = out.joints[:, config.ALL_JOINTS_TO_H36M_MAP, :][:, config.H36M_TO_J14, :]
    got = {'feat': feat, 'cam': cam, 'pose6d': pose, 'shape': shape, 'rotmats': R, 'joints': out.joints,
           'vertices': out.vertices, 'joints2d_coco': j2d, 'joints_h36mlsp': j_lsp}
    for k, v in got.items():
        assert rel_err(v.cpu().numpy(), g[k]) < RTOL, (k, rel_err(v.cpu().numpy(), g[k]))


@pytest.mark.parametrize('mode', MODES)
def test_every_block_activation_against_oracle(mode, assets_root, monkeypatch):
    """Per-layer parity (stem, pool, the 8 BasicBlock outputs and their inner activations)."""
    C, B = 17, 3
    sd = O.make_regressor_state(C, seed=7)
    reg = _regressor(C, mode, sd)
    xc = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=13))
    taps = {}
    # the default stem fuses the max pool into conv1 and never writes the stem tensor; "s2d" is the same kernel with the tensor
    # written (tests/test_gpu_stem_kernels.py holds the two bit-identical)
    monkeypatch.setenv('STRAPS_TC_CONV1', 's2d')
    with torch.no_grad():
        feat_o = O.encoder_forward(xc, sd, taps=taps)
        feat = reg.image_encoder(xc.to(DEV))
    eng = reg.image_encoder._engine
    for name, ref in taps.items():
        got = eng.read_activation(name, B).cpu().numpy()
        assert got.shape == tuple(ref.shape)
        assert rel_err(got, ref.numpy()) < RTOL, (name, rel_err(got, ref.numpy()))
    assert rel_err(feat.cpu().numpy(), feat_o.numpy()) < RTOL


@pytest.mark.parametrize('B', [1, 7, 8, 9, 64])
def test_ief_against_oracle(B, assets_root, additional_dir):
    sd = O.make_regressor_state(17, seed=3)
    reg = _regressor(17, 'fp32_simt', sd)
    feat = torch.from_numpy(np.random.RandomState(B).normal(0, 1.5, (B, 512)).astype(np.float32)).abs()
    init = O.load_initial_params(os.path.join(additional_dir, 'neutral_smpl_mean_params_6dpose.npz'))
    trace = []
    with torch.no_grad():
        p_o = O.ief_forward(feat, sd, init, 3, trace=trace)
        cam, pose, shape = reg.ief_module(feat.to(DEV))
    p = torch.cat([cam, pose, shape], 1).cpu()
    assert rel_err(p.numpy(), p_o.numpy()) < 1e-5
    reg.ief_module.iterations = 1
    with torch.no_grad():
        cam1, pose1, shape1 = reg.ief_module(feat.to(DEV))
    assert rel_err(torch.cat([cam1, pose1, shape1], 1).cpu().numpy(), trace[0].numpy()) < 1e-5


@pytest.mark.parametrize('mode', MODES)
def test_config2_full_size_b64(mode, assets_root, additional_dir, smpl_oracle):
    """BASELINE config 2 at its full size (B=64, 256x256x17): every output vs the CPU oracle."""
    import config
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat
    C, B = 17, 64
    sd = O.make_regressor_state(C, seed=WEIGHT_SEED)
    reg = _regressor(C, mode, sd)
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(DEV)
    xc = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=5))
    init = O.load_initial_params(os.path.join(additional_dir, 'neutral_smpl_mean_params_6dpose.npz'))
    with torch.no_grad():
        o = O.regress_and_pose(xc, sd, init, smpl_oracle)
        cam, pose, shape = reg(xc.to(DEV))
        R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
    for k, v in (('cam', cam), ('pose6d', pose), ('shape', shape), ('rotmats', R), ('vertices', out.vertices),
                 ('joints', out.joints)):
        e = rel_err(v.cpu().numpy(), o[k].numpy())
        assert e < RTOL, (k, e)
    # weights changed in place -> the packed copy is refreshed (version counters)
    with torch.no_grad():
        reg.ief_module.fc3.bias.add_(0.25)
        cam2, _, _ = reg(xc.to(DEV))
    assert (cam2 - cam).abs().min() > 0.1


def test_two_conv_modes_agree(assets_root):
    C, B = 18, 4
    sd = O.make_regressor_state(C, seed=9)
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=1)).to(DEV)
    with torch.no_grad():
        f0 = _regressor(C, 'fp32_simt', sd).image_encoder(x)
        f1 = _regressor(C, 'f16x3_tc', sd).image_encoder(x)
    assert rel_l2(f1.cpu().numpy(), f0.cpu().numpy()) < 5e-5   # tensor-core fp32 accumulation truncates


def test_cuda_graph_replay_is_bit_identical(assets_root):
    """The whole hot path (encoder -> IEF -> rot6d -> SMPL) captured once and replayed on new inputs equals the eager calls
    bit for bit (same kernels, same order); wrong shapes are refused."""
    import config
    from models.smpl_official import SMPL
    from straps_b200.graphs import GraphedCallable
    from utils.rigid_transform_utils import rot6d_to_rotmat
    from straps_b200._lib import StrapsError
    C, B = 17, 4
    reg = _regressor(C, 'f16x3_tc', O.make_regressor_state(C, seed=5))
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(DEV)

    def hot_path(x):
        cam, pose, shape = reg(x)
        R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
        return cam, out.vertices, out.joints
    xs = [torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=s)).to(DEV) for s in (11, 12)]
    with torch.no_grad():
        eager = [[t.clone() for t in hot_path(x)] for x in xs]
    graphed = GraphedCallable(hot_path, xs[0])
    for x, ref in zip(xs, eager):
        got = graphed(x)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(got, ref))
    assert not torch.equal(eager[0][1], eager[1][1])
    with pytest.raises(StrapsError):
        graphed(xs[0][:2])
