"""Per-tensor gradient error report of the encoder training path against the CPU oracle's autograd (debugging aid)."""
import os
os.environ.setdefault('STRAPS_TC_CONV1', 's2d')   # these diagnostics read the stem tensor, which the default fused-pool stem never writes
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'oracle'))
sys.path.insert(0, os.path.join(REPO, 'straps-3dhumanshapepose_b200'))
sys.path.insert(0, os.path.join(REPO, 'tests'))
os.environ.setdefault('STRAPS_ASSETS_ROOT', os.path.join(REPO, 'tests', '_scratch', 'assets'))
import numpy as np
import torch
import straps_oracle as O
from straps_b200 import synthetic_assets, synthetic_inputs

synthetic_assets.write_synthetic_assets(os.environ['STRAPS_ASSETS_ROOT'], seed=0)
from models.regressor import SingleInputRegressor


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


import itertools
for (C, B), MODE in itertools.product([(17, 4), (17, 16)], os.environ.get('DIAG_MODES', 'fp32_simt,f16x3_tc').split(',')):
    sd = O.make_regressor_state(C, seed=7)
    sdg = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and 'running' not in k else v.clone()) for k, v in sd.items()}
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=13))
    g = torch.from_numpy(np.random.RandomState(2).normal(0, 1, (B, 512)).astype(np.float32))
    stats, taps = {}, {}
    feat_o = O.encoder_forward(x, sdg, train=True, stats_out=stats, taps=taps)
    for t in taps.values():
        t.retain_grad()
    (feat_o * g).sum().backward()
    reg = SingleInputRegressor(C, 18, 3, conv_mode=MODE)
    reg.load_state_dict(sd)
    reg = reg.to('cuda:0').train()
    feat = reg.image_encoder(x.cuda())
    (feat * g.cuda()).sum().backward()
    torch.cuda.synchronize()
    print('=== %s C=%d B=%d  feat fwd err %.2e' % (MODE, C, B, rel(feat.detach().cpu(), feat_o.detach())))
    eng = reg._engine
    for name, ref in taps.items():
        got = eng.read_activation(name, B).cpu().numpy()
        rn = ref.detach().numpy()
        print('  act  %-12s %.2e   zero-mask mismatches %d of %d' % (name, rel(got, rn), int(((got > 0) != (rn > 0)).sum()), rn.size))
    new = reg.state_dict()
    worst = max(rel(new['image_encoder.' + k].cpu(), v) for k, v in stats.items())
    print('  running stats worst err %.2e' % worst)
    for name, p in reg.image_encoder.named_parameters():
        ref = sdg['image_encoder.' + name].grad
        print('  grad %-32s %.2e   |ref|max %.2e' % (name, rel(p.grad.cpu(), ref), float(ref.abs().max())))
