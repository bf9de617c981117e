"""CPU emulation of round 2's tap-shifted ("halo") implicit GEMM for the stride-1 3x3 convolutions (DESIGN.md 4.2, item ii).

The planned kernel keeps activations as a zero-padded raster [B][H+2][W+2][C]; an M-tile is 128 CONSECUTIVE raster positions
p0 .. p0+127; for every 64-channel chunk ONE halo tile -- raster rows p0-(W+2)-1 .. p0+127+(W+2)+1, rows outside the tensor
zero-filled (TMA out-of-bounds fill) -- is loaded, and tap (kh, kw) is the 128-row window of that tile starting at row
kh*(W+2)+kw (a UMMA descriptor with a shifted start address: tools/shift_probe.cu shows this is exact with base_offset 0).
Border positions are computed and discarded; their stores are skipped so the zero border of the output raster survives for the
next layer.  This script replays exactly that index arithmetic in numpy and checks it against torch's conv2d, including tiles
that straddle image boundaries and the first / last tile (negative / past-the-end halo rows).

    python tools/halo_emulation.py          -> prints the max error and the byte counts per tile; exits non-zero on a mismatch
"""
import sys

import numpy as np
import torch
import torch.nn.functional as F

TILE = 128


def halo_conv3x3(x_nchw, w_oihw):
    """3x3 / stride 1 / pad 1 convolution through the padded-raster + halo-tile + tap-shift scheme."""
    B, C, H, W = x_nchw.shape
    Co = w_oihw.shape[0]
    Hp, Wp = H + 2, W + 2
    raster = np.zeros((B, Hp, Wp, C), dtype=np.float64)                       # zero-padded NHWC input raster
    raster[:, 1:-1, 1:-1, :] = np.transpose(x_nchw, (0, 2, 3, 1))
    flat = raster.reshape(B * Hp * Wp, C)
    n_pos = flat.shape[0]
    out = np.zeros((B * Hp * Wp, Co), dtype=np.float64)                       # padded OUTPUT raster, border stays zero
    wk = np.transpose(w_oihw, (2, 3, 1, 0)).astype(np.float64)                # [kh][kw][ci][co]
    halo_rows = TILE + 2 * Wp + 2
    for p0 in range(0, n_pos, TILE):
        start = p0 - Wp - 1                                                   # raster row of halo row 0 (may be negative)
        halo = np.zeros((halo_rows, C))
        lo, hi = max(start, 0), min(start + halo_rows, n_pos)
        if hi > lo:
            halo[lo - start:hi - start] = flat[lo:hi]                         # everything else: TMA zero fill
        acc = np.zeros((TILE, Co))
        for kh in range(3):
            for kw in range(3):
                shift = kh * Wp + kw                                          # descriptor start row of this tap
                acc += halo[shift:shift + TILE] @ wk[kh, kw]
        for r in range(TILE):                                                 # epilogue: one thread per tile row
            p = p0 + r
            if p >= n_pos:
                continue
            y, x = (p // Wp) % Hp, p % Wp
            if 1 <= y <= H and 1 <= x <= W:                                   # border positions are computed but never stored
                out[p] = acc[r]
    res = out.reshape(B, Hp, Wp, Co)
    assert not res[:, 0].any() and not res[:, -1].any() and not res[:, :, 0].any() and not res[:, :, -1].any()
    return np.transpose(res[:, 1:-1, 1:-1, :], (0, 3, 1, 2)), halo_rows


def main():
    rng = np.random.RandomState(0)
    worst = 0.0
    for (B, C, H, W, Co) in ((3, 8, 8, 8, 4), (2, 4, 16, 16, 8), (5, 3, 6, 10, 2), (1, 2, 64, 64, 2)):
        x = rng.normal(0, 1, (B, C, H, W))
        w = rng.normal(0, 1, (Co, C, 3, 3))
        got, halo_rows = halo_conv3x3(x, w)
        ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), padding=1).numpy()
        err = float(np.abs(got - ref).max())
        worst = max(worst, err)
        print('B=%d C=%d %dx%d -> Cout=%d: max abs error %.2e   (halo tile %d rows for a %d-position M-tile, border waste %.0f %%)'
              % (B, C, H, W, Co, err, halo_rows, TILE, 100.0 * ((H + 2) * (W + 2) / (H * W) - 1)))
    for name, hw, cin, bn in (('layer1', 64, 64, 64), ('layer2', 32, 128, 128)):
        wp = hw + 2
        a_now, a_halo = 9 * (cin // 64) * 2 * TILE * 128, (cin // 64) * 2 * (TILE + 2 * wp + 2) * 128
        w_bytes = 9 * (cin // 64) * 2 * bn * 128
        print('%s: A bytes per tile %d KB -> %d KB, with the weights %d KB -> %d KB' % (name, a_now // 1024, a_halo // 1024,
                                                                                     (a_now + w_bytes) // 1024, (a_halo + w_bytes) // 1024))
    if worst > 1e-10:
        print('MISMATCH')
        sys.exit(1)
    print('halo scheme == conv2d (max abs error %.2e)' % worst)


if __name__ == '__main__':
    main()
