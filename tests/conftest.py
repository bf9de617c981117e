import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, 'straps-3dhumanshapepose_b200')
ORACLE = os.path.join(REPO, 'oracle')
GOLDEN = os.path.join(REPO, 'tests', 'golden')
SCRATCH = os.path.join(REPO, 'tests', '_scratch')
ASSETS_ROOT = os.path.join(SCRATCH, 'assets')

# the drop-in source root and the oracle (test infrastructure) -- config.py reads STRAPS_ASSETS_ROOT at import
os.environ.setdefault('STRAPS_ASSETS_ROOT', ASSETS_ROOT)
for p in (ORACLE, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

ASSET_SEED, WEIGHT_SEED, INPUT_SEED = 0, 1, 3
RTOL = 1e-4   # north_star: 1e-4 relative fp32 (max-abs error / max-abs reference, per tensor)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """`python -m pytest tests` on a machine without CUDA skips the gpu-marked tests instead of erroring in their fixtures."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='needs a CUDA device (B200): run with -m gpu on the GPU box')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def rel_err(a, b):
    """max-abs error relative to the max-abs of the reference tensor (SURVEY.md 8c tolerance definition)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope='session')
def assets_root():
    from straps_b200 import synthetic_assets
    synthetic_assets.write_synthetic_assets(ASSETS_ROOT, seed=ASSET_SEED)
    return ASSETS_ROOT


@pytest.fixture(scope='session')
def additional_dir(assets_root):
    return os.path.join(assets_root, 'additional')


@pytest.fixture(scope='session')
def smpl_oracle(additional_dir):
    import straps_oracle as O
    return O.SmplOracle(additional_dir, batch_size=1)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def smpl_inputs(batch, seed):
    """Same generator as oracle/gen_golden.py:smpl_inputs."""
    rng = np.random.RandomState(seed)
    betas = rng.normal(0, 1, (batch, 10)).astype(np.float32)
    pose6d = rng.normal(0, 1, (batch, 144)).astype(np.float32)
    aa = rng.normal(0, 0.4, (batch, 72)).astype(np.float32)
    return betas, pose6d, aa


def checksum(a):
    return float(np.asarray(a, dtype=np.float64).sum())
