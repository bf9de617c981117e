"""GPU: the 3-pass fp16-split convolutions outside the benign statistics of the other tests (round-1 review, weak #2).

make_regressor_state draws BatchNorm gamma, sigma^2 in [0.6, 1.4]; real checkpoints are not that polite.  Here
  * every convolution's output channels are rescaled by 10^U(-1.5, 1.5), so the running variances span 1e-3 .. 1e3 and the folded
    BatchNorm scale gamma / sigma spans six decades inside a layer (the per-row power-of-two weight scale has to absorb it);
  * gamma itself spans three decades per layer around a per-layer target magnitude: a near-zero layer (activations peak at ~4e-3: the
    low halves of an unscaled split would be fp16-subnormal), layers near the top of the fp16 range (peaks of 3e4 .. 6e4, the clamp is
    at 65504), and ordinary ones in between;
  * the running statistics are the batch statistics of a calibration pass, so eval-mode activations really have those magnitudes.
Every block activation and the features are held to the north-star bar (1e-4 of the tensor maximum) against the CPU oracle, and a
plain C = 18 run at the bench batch (B = 64) closes the other gap named there."""
import os

import numpy as np
import pytest
import torch

import straps_oracle as O
from conftest import rel_err, RTOL
from straps_b200 import synthetic_inputs

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

# per-layer target magnitude of gamma (the post-BatchNorm standard deviation of the strongest channels)
_TARGETS = {'bn1': 3.0, 'layer1.0.bn1': 1e-3, 'layer1.0.bn2': 0.5, 'layer1.1.bn1': 7e3, 'layer1.1.bn2': 2.0,
            'layer2.0.bn1': 1.0, 'layer2.0.bn2': 6e3, 'layer2.0.downsample.1': 5e3, 'layer2.1.bn1': 2e-3, 'layer2.1.bn2': 3e3,
            'layer3.0.bn1': 1.0, 'layer3.0.bn2': 1e-2, 'layer3.0.downsample.1': 2e-2, 'layer3.1.bn1': 7e3, 'layer3.1.bn2': 1e-2,
            'layer4.0.bn1': 30.0, 'layer4.0.bn2': 1.0, 'layer4.0.downsample.1': 0.5, 'layer4.1.bn1': 1e-3, 'layer4.1.bn2': 4.0}


def adversarial_state(C, seed, x_cal):
    rng = np.random.RandomState(seed)
    sd = O.make_regressor_state(C, seed=seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.float32)))
    for conv, bn, cin, cout, k, stride, pad in O.conv_bn_names():
        s = 10.0 ** rng.uniform(-1.5, 1.5, cout)
        sd['image_encoder.' + conv + '.weight'] = sd['image_encoder.' + conv + '.weight'] * t(s).view(-1, 1, 1, 1)
        gamma = _TARGETS[bn] * 10.0 ** rng.uniform(-3, 0, cout)
        gamma[rng.randint(cout)] = _TARGETS[bn]                              # at least one channel at the target
        sd['image_encoder.' + bn + '.weight'] = t(gamma * rng.choice([-1.0, 1.0], cout))
        sd['image_encoder.' + bn + '.bias'] = t(gamma * rng.normal(0, 0.3, cout))
        sd['image_encoder.' + bn + '.running_mean'] = torch.zeros(cout)
        sd['image_encoder.' + bn + '.running_var'] = torch.zeros(cout)
    # calibration: one train-mode pass of the oracle; with zero initial running statistics and momentum 0.1 the batch statistics are
    # 10 x the updated running ones
    stats = {}
    with torch.no_grad():
        O.encoder_forward(x_cal, sd, train=True, stats_out=stats)
    for k, v in stats.items():
        sd['image_encoder.' + k] = (v * 10.0).contiguous()
    return sd


def test_adversarial_batchnorm_statistics_and_activation_range(assets_root):
    from models.regressor import SingleInputRegressor
    C, B = 18, 8
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=77))
    sd = adversarial_state(C, 5, x)
    var = torch.cat([sd['image_encoder.%s.running_var' % bn] for _, bn, *_ in O.conv_bn_names()])
    assert float(var.max() / var.min()) > 1e5                               # the statistics really are spread out
    taps = {}
    with torch.no_grad():
        feat_o = O.encoder_forward(x, sd, taps=taps)
    peaks = {k: float(v.abs().max()) for k, v in taps.items()}
    assert min(peaks.values()) < 2e-2 and 2e4 < max(peaks.values()) < 65504, peaks       # a near-zero layer and one near the clamp
    reg = SingleInputRegressor(C, 18, 3)
    reg.load_state_dict(sd)
    reg = reg.to(DEV).eval()
    os.environ['STRAPS_TC_CONV1'] = 's2d'                                   # so that the stem tensor can be read back
    try:
        with torch.no_grad():
            feat = reg.image_encoder(x.to(DEV))
        errs = {name: rel_err(reg._engine.read_activation(name, B).cpu().numpy(), ref.numpy()) for name, ref in taps.items()}
    finally:
        os.environ.pop('STRAPS_TC_CONV1', None)
    errs['feat'] = rel_err(feat.cpu().numpy(), feat_o.numpy())
    print('adversarial range: peaks %s\nerrors %s' % ({k: '%.1e' % v for k, v in peaks.items()}, {k: '%.1e' % v for k, v in errs.items()}))
    assert all(e < RTOL for e in errs.values()), errs
    with torch.no_grad():                                                   # the default (fused-pool) stem on the same weights
        feat2 = reg.image_encoder(x.to(DEV))
    assert torch.equal(feat, feat2)


def test_c18_at_the_bench_batch(assets_root, additional_dir, smpl_oracle):
    """C = 18 (run_train.py:35) at B = 64, whole path: encoder -> IEF -> rot6d -> SMPL against the oracle."""
    import config
    from models.regressor import SingleInputRegressor
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat
    C, B = 18, 64
    sd = O.make_regressor_state(C, seed=2)
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=8))
    init = O.load_initial_params(os.path.join(additional_dir, 'neutral_smpl_mean_params_6dpose.npz'))
    so = O.SmplOracle(additional_dir, batch_size=B)
    with torch.no_grad():
        o = O.regress_and_pose(x, sd, init, so)
    reg = SingleInputRegressor(C, 18, 3)
    reg.load_state_dict(sd)
    reg = reg.to(DEV).eval()
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(DEV)
    with torch.no_grad():
        cam, pose, shape = reg(x.to(DEV))
        R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
    for k, v in (('cam', cam), ('pose6d', pose), ('shape', shape), ('vertices', out.vertices), ('joints', out.joints)):
        assert rel_err(v.cpu().numpy(), o[k].numpy()) < RTOL, (k, rel_err(v.cpu().numpy(), o[k].numpy()))
