"""CPU emulation of round 2's tap-shifted ("halo") implicit GEMM for the stride-1 3x3 convolutions (DESIGN.md 4.2, item ii;
the device code is csrc/conv_tc.cu: conv_halo_kernel).

Activations stay UNPADDED NHWC in HBM.  Per image the outputs are enumerated on a virtual zero-padded raster of (H+2) x (W+2)
positions; an M-tile is 128 CONSECUTIVE raster positions p0 .. p0+127 of one image.  For every 64-channel chunk ONE halo tile is
loaded with a single TMA box {64 channels, W+2 pixels, RH rows} that starts at pixel -1 / row r0-1 of the image: the out-of-bounds
fill of TMA materialises the zero border, so shared memory holds padded rows r0 .. r0+RH-1 of the raster, one 128-byte line per
position.  Tap (kh, kw) of the 3x3 filter is then the 128-line window that starts at line (p0 - r0*(W+2)) + (kh-1)*(W+2) + (kw-1)
-- a UMMA descriptor with a shifted start address (tools/shift_probe.cu: exact with base_offset 0).  Border positions are computed
and discarded.  This script replays exactly that index arithmetic in numpy and checks it against torch's conv2d, and checks the
static bounds the kernel relies on (rows per box, window inside the tile, last tile of an image empty).

    python tools/halo_emulation.py          -> prints the max error and the byte counts per tile; exits non-zero on a mismatch
"""
import sys

import numpy as np
import torch
import torch.nn.functional as F

TILE = 128


def halo_rows(W):
    """Rows of the TMA box: the 128 + 2(W+2) + 2 consecutive positions a tile touches span at most this many padded rows."""
    wp = W + 2
    return -(-(TILE + 2 * wp + 2) // wp) + 1


def tiles_per_image(H, W):
    """Tiles with at least one interior position: the raster's last row is all border, so a tile that starts in it is skipped."""
    hp, wp = H + 2, W + 2
    last_interior = H * wp + W                       # raster index of interior position (H, W)
    return last_interior // TILE + 1


def halo_conv3x3(x_nchw, w_oihw, mt=1):
    """3x3 / stride 1 / pad 1 convolution through the virtual padded raster + halo box + tap-shift scheme.
    mt = 2: conv_halo2_kernel -- two consecutive tiles per work item share one (taller) box."""
    if mt > 1:
        return halo_conv3x3_items(x_nchw, w_oihw, mt)
    B, C, H, W = x_nchw.shape
    Co = w_oihw.shape[0]
    Hp, Wp = H + 2, W + 2
    RH = halo_rows(W)
    x_nhwc = np.transpose(x_nchw, (0, 2, 3, 1)).astype(np.float64)
    out = np.zeros((B, H, W, Co), dtype=np.float64)                           # unpadded output
    wk = np.transpose(w_oihw, (2, 3, 1, 0)).astype(np.float64)                # [kh][kw][ci][co]
    for b in range(B):
        for t in range(tiles_per_image(H, W)):
            p0 = t * TILE
            r0 = (p0 - Wp - 1) // Wp                                          # floor division: -1 for the first tile
            # the TMA box {C, Wp, RH} at (x = -1, y = r0 - 1): out-of-range pixels / rows arrive as zeros
            halo = np.zeros((RH, Wp, C))
            for i in range(RH):
                y = r0 - 1 + i                                                # image row of padded row r0 + i
                if 0 <= y < H:
                    halo[i, 1:W + 1] = x_nhwc[b, y]
            lines = halo.reshape(RH * Wp, C)                                  # one 128-byte line per raster position
            base = p0 - r0 * Wp                                               # line of the tile's first position
            acc = np.zeros((TILE, Co))
            for kh in range(3):
                for kw in range(3):
                    start = base + (kh - 1) * Wp + (kw - 1)                   # descriptor start line of this tap
                    assert 0 <= start and start + TILE <= RH * Wp, (p0, kh, kw, start, RH * Wp)
                    acc += lines[start:start + TILE] @ wk[kh, kw]
            for r in range(TILE):                                             # epilogue: one thread per tile row
                p = p0 + r
                py, px = p // Wp, p % Wp
                if 1 <= py <= H and 1 <= px <= W:                             # border / past-the-end positions are never stored
                    out[b, py - 1, px - 1] = acc[r]
    return np.transpose(out, (0, 3, 1, 2)), RH


def halo_conv3x3_items(x_nchw, w_oihw, mt):
    """conv_halo2_kernel's arithmetic: an item = mt consecutive tiles; one box of rows(mt) raster rows; tile t of the item reads the
    windows that start t * 128 lines further down."""
    B, C, H, W = x_nchw.shape
    Co = w_oihw.shape[0]
    Wp = W + 2
    RH = -(-(mt * TILE + 2 * Wp + 2) // Wp) + 1
    items = -(-tiles_per_image(H, W) // mt)
    x_nhwc = np.transpose(x_nchw, (0, 2, 3, 1)).astype(np.float64)
    out = np.zeros((B, H, W, Co), dtype=np.float64)
    wk = np.transpose(w_oihw, (2, 3, 1, 0)).astype(np.float64)
    for b in range(B):
        for it in range(items):
            p0 = it * mt * TILE
            r0 = (p0 + Wp - 1) // Wp - 2
            halo = np.zeros((RH, Wp, C))
            for i in range(RH):
                y = r0 - 1 + i
                if 0 <= y < H:
                    halo[i, 1:W + 1] = x_nhwc[b, y]
            lines = halo.reshape(RH * Wp, C)
            base = p0 - r0 * Wp
            for t in range(mt):
                acc = np.zeros((TILE, Co))
                start = base - Wp - 1
                for kh in range(3):
                    for kw in range(3):
                        s0 = start + t * TILE
                        assert 0 <= s0 and s0 + TILE <= RH * Wp, (p0, t, kh, kw, s0, RH * Wp)
                        acc += lines[s0:s0 + TILE] @ wk[kh, kw]
                        start += 1
                    start += Wp - 3
                for r in range(TILE):
                    p = p0 + t * TILE + r
                    py, px = p // Wp, p % Wp
                    if 1 <= py <= H and 1 <= px <= W:
                        out[b, py - 1, px - 1] = acc[r]
    return np.transpose(out, (0, 3, 1, 2)), RH


def main():
    rng = np.random.RandomState(0)
    worst = 0.0
    for (B, C, H, W, Co) in ((3, 8, 8, 8, 4), (2, 4, 16, 16, 8), (2, 3, 32, 32, 2), (1, 2, 64, 64, 2)):
        x = rng.normal(0, 1, (B, C, H, W))
        w = rng.normal(0, 1, (Co, C, 3, 3))
        got, RH = halo_conv3x3(x, w)
        ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), padding=1).numpy()
        err = float(np.abs(got - ref).max())
        worst = max(worst, err)
        print('B=%d C=%d %dx%d -> Cout=%d: max abs error %.2e   (box of %d rows x %d pixels, %d tiles per image instead of %.1f)'
              % (B, C, H, W, Co, err, RH, W + 2, tiles_per_image(H, W), H * W / TILE))
    x = rng.normal(0, 1, (2, 2, 64, 64))
    w = rng.normal(0, 1, (2, 2, 3, 3))
    got, RH = halo_conv3x3(x, w, mt=2)
    ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), padding=1).numpy()
    err = float(np.abs(got - ref).max())
    worst = max(worst, err)
    print('two tiles per item (conv_halo2_kernel), 64x64: max abs error %.2e   (box of %d rows, %d items per image)'
          % (err, RH, -(-tiles_per_image(64, 64) // 2)))
    assert RH == 7
    for name, hw, cin, bn in (('layer1', 64, 64, 64), ('layer2', 32, 128, 128)):
        wp, rh = hw + 2, halo_rows(hw)
        a_now, a_halo = 9 * (cin // 64) * 2 * TILE * 128, (cin // 64) * 2 * rh * wp * 128
        w_bytes = 9 * (cin // 64) * 2 * bn * 128
        n_now, n_halo = hw * hw // TILE, tiles_per_image(hw, hw)
        print('%s: box %d x %d lines; bytes per image %d KB -> %d KB (A %d -> %d KB per tile, %d -> %d tiles)'
              % (name, rh, wp, n_now * (a_now + w_bytes) // 1024, n_halo * (a_halo + w_bytes) // 1024, a_now // 1024, a_halo // 1024,
                 n_now, n_halo))
    if worst > 1e-10:
        print('MISMATCH')
        sys.exit(1)
    print('halo scheme == conv2d (max abs error %.2e)' % worst)


if __name__ == '__main__':
    main()
