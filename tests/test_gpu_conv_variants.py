"""GPU: the experimental scheduling variants of conv_tc_kernel (environment switches read once per process, hence one
subprocess per variant) produce the same encoder features as the shipped configuration -- they only change how tiles are
staged (cluster weight multicast, half-size SWIZZLE_64B stages, wide tiles), never the arithmetic order within a tile.
Tolerance: bit-identical features where the per-tile MMA order is unchanged, else the 1e-4 bar against the fp32 kernels."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO, PKG, ORACLE

pytestmark = pytest.mark.gpu

_SCRIPT = r'''
import json, sys, os
sys.path.insert(0, %(pkg)r); sys.path.insert(0, %(oracle)r)
os.environ.setdefault('STRAPS_ASSETS_ROOT', %(assets)r)
import numpy as np, torch
from straps_b200 import synthetic_assets, synthetic_inputs
synthetic_assets.write_synthetic_assets(os.environ['STRAPS_ASSETS_ROOT'], seed=0)
import straps_oracle as O
from models.regressor import SingleInputRegressor
out = {}
for C, B in ((17, 3), (18, 5)):          # odd tile counts: the last cluster has an out-of-range M-tile
    sd = O.make_regressor_state(C, seed=9)
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=1)).cuda()
    feats = {}
    for mode in ('fp32_simt', 'f16x3_tc'):
        reg = SingleInputRegressor(C, 18, 3, conv_mode=mode)
        reg.load_state_dict(sd)
        reg = reg.cuda().eval()
        with torch.no_grad():
            feats[mode] = reg.image_encoder(x).double().cpu().numpy()
    a, b = feats['f16x3_tc'], feats['fp32_simt']
    out['C%%d' %% C] = {'rel_l2_vs_fp32': float(np.linalg.norm(a - b) / np.linalg.norm(b)), 'checksum': float(a.sum()),
                      'abs_checksum': float(np.abs(a).sum())}
print(json.dumps(out))
'''


def _run(env_extra):
    env = dict(os.environ)
    for k in ('STRAPS_TC_MCAST', 'STRAPS_TC_BK', 'STRAPS_TC_TILES', 'STRAPS_TC_PAIR', 'STRAPS_TC_DS_OVERLAP', 'STRAPS_TC_EPI_WARPS'):
        env.pop(k, None)
    env.update(env_extra)
    script = _SCRIPT % {'pkg': PKG, 'oracle': ORACLE, 'assets': os.path.join(REPO, 'tests', '_scratch', 'assets')}
    res = subprocess.run([sys.executable, '-c', script], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


@pytest.fixture(scope='module')
def shipped(assets_root):
    return _run({})


# variants written after the round's GPU budget was spent: compiled and reviewed but never run on hardware; they join the list above
# once a run with STRAPS_TEST_UNVERIFIED=1 has passed
_UNVERIFIED = pytest.mark.skipif(not os.environ.get('STRAPS_TEST_UNVERIFIED'), reason='not yet run on hardware (set STRAPS_TEST_UNVERIFIED=1)')


@pytest.mark.parametrize('env', [{'STRAPS_TC_MCAST': '2'}, {'STRAPS_TC_MCAST': '4'}, {'STRAPS_TC_BK': '32'},
                                 {'STRAPS_TC_BK': '32', 'STRAPS_TC_MCAST': '2'}, {'STRAPS_TC_DS_OVERLAP': '0'},
                                 pytest.param({'STRAPS_TC_EPI_WARPS': '8'}, marks=_UNVERIFIED)],
                         ids=lambda e: ','.join('%s=%s' % (k[10:], v) for k, v in e.items()))
def test_variant_matches_shipped_configuration(env, shipped):
    got = _run(env)
    for key, ref in shipped.items():
        assert ref['rel_l2_vs_fp32'] < 5e-5, ref            # the shipped configuration itself
        assert got[key]['rel_l2_vs_fp32'] < 5e-5, (env, got[key])
        # same MMAs in the same order into the same accumulators -> identical bits (BK = 32 splits a K-block in two stages but
        # keeps the K order)
        assert got[key]['checksum'] == ref['checksum'] and got[key]['abs_checksum'] == ref['abs_checksum'], (env, got[key], ref)
