"""GPU: the three stem implementations -- conv1 from the pixel-pair layout with the max pool fused into its epilogue (default,
STRAPS_TC_CONV1 unset / "s2dp"), the same kernel writing the stem tensor + the stand-alone pool ("s2d"), and the round-1 im2col form
through conv_tc_kernel ("im2col") -- on the same weights and input.  All three issue the same MMAs in the same order per output pixel
and max is exact, so the stem, the pooled planes and the features must be BIT-IDENTICAL (reference models/resnet.py:202-206);
the pooled planes must also equal torch's max_pool2d of the stem bit for bit.  Batches 3 and 5 make the contiguous row ranges of
the fused-pool kernel start inside images and leave CTAs with short ranges; B = 64 is the bench configuration."""
import os

import numpy as np
import pytest
import torch

import straps_oracle as O
from conftest import rel_err, RTOL
from straps_b200 import synthetic_inputs

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _run(reg, x, mode, names):
    old = os.environ.get('STRAPS_TC_CONV1')
    if mode is None:
        os.environ.pop('STRAPS_TC_CONV1', None)
    else:
        os.environ['STRAPS_TC_CONV1'] = mode
    try:
        with torch.no_grad():
            feat = reg.image_encoder(x)
        acts = {n: reg._engine.read_activation(n, x.shape[0]).clone() for n in names}
        return feat.clone(), acts
    finally:
        if old is None:
            os.environ.pop('STRAPS_TC_CONV1', None)
        else:
            os.environ['STRAPS_TC_CONV1'] = old


@pytest.mark.parametrize('C,B', [(17, 3), (18, 5), (17, 64)])
def test_stem_kernels_are_bit_identical(C, B, assets_root):
    from models.regressor import SingleInputRegressor
    from straps_b200._lib import StrapsError
    sd = O.make_regressor_state(C, seed=9)
    reg = SingleInputRegressor(C, 18, 3)
    reg.load_state_dict(sd)
    reg = reg.to(DEV).eval()
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=1)).to(DEV)
    f_fused, a_fused = _run(reg, x, None, ['pool', 'layer1.0'])
    with pytest.raises(StrapsError):                  # the fused kernel never writes the stem tensor
        reg._engine.read_activation('stem', B)
    f_s2d, a_s2d = _run(reg, x, 's2d', ['stem', 'pool', 'layer1.0'])
    f_old, a_old = _run(reg, x, 'im2col', ['stem', 'pool', 'layer1.0'])
    assert torch.equal(a_s2d['stem'], a_old['stem'])
    assert torch.equal(a_s2d['pool'], a_old['pool'])
    assert torch.equal(a_fused['pool'], a_s2d['pool'])
    assert torch.equal(a_fused['layer1.0'], a_old['layer1.0'])
    assert torch.equal(f_fused, f_old) and torch.equal(f_s2d, f_old)
    # the pool itself: exact max over the 3x3 / stride 2 / pad 1 window of the stem (fp16 hi + lo re-summed exactly in fp32)
    ref_pool = torch.nn.functional.max_pool2d(a_s2d['stem'], 3, 2, 1)
    assert rel_err(a_fused['pool'].cpu().numpy(), ref_pool.cpu().numpy()) < 1e-6
    # and against the CPU oracle (small batches only: the oracle's fp32 convolutions take seconds per image)
    if B <= 5:
        taps = {}
        with torch.no_grad():
            fo = O.encoder_forward(x.cpu(), sd, taps=taps)
        assert rel_err(a_fused['pool'].cpu().numpy(), taps['pool'].numpy()) < RTOL
        assert rel_err(f_fused.cpu().numpy(), fo.numpy()) < RTOL


@pytest.mark.parametrize('C,B', [(18, 4), (17, 64)])
def test_forward_from_labels_is_bit_identical_to_the_assembled_input(C, B, assets_root):
    """SingleInputRegressor.forward_from_labels (proxy representation generated inside the stem's input pack) against the reference's
    own assembly x = cat([binary(seg), heatmaps(joints2D)]) (train/train_synthetic_otf_rendering.py:178-182) through the same
    regressor: same planes, hence the same bits.  Joints straddle the borders (clipped windows) and leave the image (empty maps)."""
    from models.regressor import SingleInputRegressor
    from utils.label_conversions import convert_2Djoints_to_gaussian_heatmaps_torch, convert_multiclass_to_binary_labels_torch
    sd = O.make_regressor_state(C, seed=4)
    reg = SingleInputRegressor(C, 18, 3)
    reg.load_state_dict(sd)
    reg = reg.to(DEV).eval()
    rng = np.random.RandomState(B)
    seg = torch.from_numpy(rng.randint(0, 7, (B, 256, 256)).astype(np.float32) * (rng.uniform(0, 1, (B, 256, 256)) < 0.3)).float().to(DEV)
    j2d = torch.from_numpy(rng.uniform(-30, 290, (B, C - 1, 2)).astype(np.float32)).to(DEV)
    j2d[0, 0] = torch.tensor([0.0, 0.0])
    j2d[0, 1] = torch.tensor([255.9, 255.9])
    j2d[0, 2] = torch.tensor([-7.5, 100.0])
    j2d[0, 3] = torch.tensor([100.0, 262.9])
    with torch.no_grad():
        x = torch.cat([convert_multiclass_to_binary_labels_torch(seg).unsqueeze(1),
                       convert_2Djoints_to_gaussian_heatmaps_torch(j2d, 256)], dim=1)
        cam, pose, shape = [t.clone() for t in reg(x)]
        pool = reg._engine.read_activation('pool', B).clone()
        cam2, pose2, shape2 = reg.forward_from_labels(seg, j2d)
        pool2 = reg._engine.read_activation('pool', B)
    assert torch.equal(pool, pool2)
    assert torch.equal(cam, cam2) and torch.equal(pose, pose2) and torch.equal(shape, shape2)
    reg.train()
    with pytest.raises(RuntimeError):
        reg.forward_from_labels(seg, j2d)
