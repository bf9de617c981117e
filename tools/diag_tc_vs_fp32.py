"""Per-tensor difference between the tensor-core and the fp32 CUDA-core encoder backward on the SAME saved activations."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'oracle'))
sys.path.insert(0, os.path.join(REPO, 'straps-3dhumanshapepose_b200'))
os.environ.setdefault('STRAPS_ASSETS_ROOT', os.path.join(REPO, 'tests', '_scratch', 'assets'))
import numpy as np
import torch
import straps_oracle as O
from straps_b200 import synthetic_assets, synthetic_inputs

synthetic_assets.write_synthetic_assets(os.environ['STRAPS_ASSETS_ROOT'], seed=0)
from models.regressor import SingleInputRegressor


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


for C, B in [(17, int(b)) for b in os.environ.get('DIAG_B', '4,64').split(',')]:
    sd = O.make_regressor_state(C, seed=11)
    reg = SingleInputRegressor(C, 18, 3, conv_mode='f16x3_tc')
    reg.load_state_dict(sd)
    reg = reg.to('cuda:0').train()
    dev = torch.device('cuda:0')
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=5)).to(dev)
    g = torch.from_numpy(np.random.RandomState(3).normal(0, 1, (B, 512)).astype(np.float32)).to(dev)
    eng = reg._engine
    h = eng._sync(dev, B, C)
    conv_w, bn, _, _ = eng._train_tensors(dev)
    shapes, chans = [tuple(w.shape) for w in conv_w], [q[0].shape[0] for q in bn]
    h.encoder_train_forward(x, update_running_stats=False, mode='f16x3_tc')
    def truth_conv1():
        """fp64 weight gradient of conv1 from the dY the last backward left behind (torch, GPU)."""
        dy = h.read_activation('grad:conv1', B).double()
        return torch.nn.grad.conv2d_weight(x.double(), shapes[0], dy, stride=2, padding=3)

    a = h.encoder_backward(g, shapes, chans)
    t_a = truth_conv1()
    a2 = h.encoder_backward(g, shapes, chans)
    os.environ['STRAPS_WGRAD'] = 'simt'            # tensor-core data gradients, fp32 CUDA-core weight gradients: same dY for both
    hyb = h.encoder_backward(g, shapes, chans)
    del os.environ['STRAPS_WGRAD']
    b = h.encoder_backward(g, shapes, chans, mode='fp32_simt')
    t_b = truth_conv1()
    print('conv1 dW against the fp64 weight gradient of the SAME dY:  tensor-core %.2e   fp32 CUDA-core %.2e   (the two dY differ by %.2e)'
          % (rel(a[0][0], t_a), rel(b[0][0], t_b), rel(t_a, t_b)))
    b2 = h.encoder_backward(g, shapes, chans, mode='fp32_simt')
    print('=== C=%d B=%d   columns: tc vs fp32 | tc run-to-run | fp32 run-to-run | tc vs (tc dgrad + fp32 wgrad)' % (C, B))
    for i in range(20):
        print('  conv %2d %-18s dW %.2e | %.2e | %.2e | %.2e    dgamma %.2e  dbeta %.2e' % (
            i, str(shapes[i]), rel(a[0][i], b[0][i]), rel(a[0][i], a2[0][i]), rel(b[0][i], b2[0][i]), rel(a[0][i], hyb[0][i]),
            rel(a[1][i][0], b[1][i][0]), rel(a[1][i][1], b[1][i][1])))
