"""Per-layer error report of the encoder (both conv modes) against the CPU oracle -- a debugging aid.
Usage: python tools/diag_encoder.py [mode ...]    (run on a GPU box; prints one line per activation)"""
import os
os.environ.setdefault('STRAPS_TC_CONV1', 's2d')   # these diagnostics read the stem tensor, which the default fused-pool stem never writes
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

modes = sys.argv[1:] or ['fp32_simt', 'f16x3_tc']
C, B = 17, 3
sd = O.make_regressor_state(C, seed=7)
xc = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=13))
taps = {}
with torch.no_grad():
    feat_o = O.encoder_forward(xc, sd, taps=taps)
for mode in modes:
    reg = SingleInputRegressor(C, 18, 3, conv_mode=mode)
    reg.load_state_dict(sd)
    reg = reg.to('cuda:0').eval()
    try:
        with torch.no_grad():
            feat = reg.image_encoder(xc.cuda())
        torch.cuda.synchronize()
    except Exception as e:
        print(mode, 'FAILED:', e)
        continue
    eng = reg.image_encoder._engine
    for name, ref in taps.items():
        got = eng.read_activation(name, B).cpu().numpy()
        r = ref.numpy()
        err = np.abs(got - r).max() / np.abs(r).max()
        print('%-10s %-12s rel_err %.3e  |ref|max %.3e  got max %.3e  nan %d' % (mode, name, err, np.abs(r).max(), np.abs(got).max(), int(np.isnan(got).sum())))
    print('%-10s %-12s rel_err %.3e' % (mode, 'feat', (feat.cpu() - feat_o).abs().max() / feat_o.abs().max()))
