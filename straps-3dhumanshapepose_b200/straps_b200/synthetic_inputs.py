"""Seeded synthetic proxy-representation inputs (silhouette + 2-D joint heatmaps).

Mirrors the input distribution the reference feeds the regressor
(train/train_synthetic_otf_rendering.py:178-182, utils/label_conversions.py:90-127): channel 0 is a
binary silhouette, every further channel holds one 16x16 truncated Gaussian (sigma 4, peak 0.982 on
the linspace(-8, 8, 16) grid) pasted at an integer joint centre.  numpy RandomState only, so the
same seed gives the same bytes on every machine.
"""
import numpy as np

IMG_WH = 256


def _gaussian16():
    g = np.linspace(-8.0, 8.0, 16, dtype=np.float32)
    d2 = g[:, None] ** 2 + g[None, :] ** 2
    return np.exp(-(d2 / (2.0 * 4.0 ** 2))).astype(np.float32)


def make_proxy_batch(batch, channels=17, seed=0, out=None):
    """-> float32 ndarray [batch, channels, 256, 256]."""
    rng = np.random.RandomState(seed)
    x = out if out is not None else np.zeros((batch, channels, IMG_WH, IMG_WH), dtype=np.float32)
    x[...] = 0
    g = _gaussian16()
    yy, xx = np.mgrid[0:IMG_WH, 0:IMG_WH]
    for b in range(batch):
        cy, cx = rng.uniform(100, 156, 2)
        ry, rx = rng.uniform(70, 110), rng.uniform(35, 60)
        x[b, 0] = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0).astype(np.float32)
        for j in range(1, channels):
            jy, jx = rng.randint(8, 248, 2)
            x[b, j, jy - 8:jy + 8, jx - 8:jx + 8] = g
    return x
