"""CPU emulation of conv1_s2d_kernel's index arithmetic (csrc/conv_tc.cu): the 7x7 / stride 2 / pad 3 stem convolution computed from
the pixel-PAIR layout of the padded input, with tap pairs as line-shifted windows of one TMA box per filter row.

Replayed exactly as the device code does it:
  pack_input_s2d_kernel   xs[b][h + 3][(w + 3) >> 1][((w + 3) & 1) * 24 + c] = x[b][c][h][w]      (262 rows x 132 pairs x PITCH)
  pack_w_tc_kernel        wrow[co][kh * 192 + kw * 24 + c] = w[co][c][kh][kw]   (kw = 7 and c >= C are zero)    -- conv1 branch
  producer                box {PITCH, 131 pairs} of padded row 2 oh + kh  ->  131 shared-memory lines of 64 elements
  MMA issuer              K step j = 0..10 of filter row kh: A = lines [j // 3, j // 3 + 128), elements [(j % 3) * 16, +16);
                          W = wrow[:, kh * 192 + 16 j : +16]   (weight chunk j // 4, byte offset (j % 4) * 32)
and compared with torch's conv2d; the fused max-pool epilogue is replayed per CTA as well.

    python tools/conv1_s2d_emulation.py      -> prints the max error; exits non-zero on a mismatch
"""
import sys

import numpy as np
import torch
import torch.nn.functional as F

XS_H, XS_PAIRS, LINES, XP_C, KROW = 262, 132, 131, 24, 192


def pack_input(x, pitch):
    B, C, H, W = x.shape
    assert H == 256 and W == 256 and C <= XP_C
    xs = np.zeros((B, XS_H, XS_PAIRS, pitch))
    for w in range(W):
        pw = w + 3
        xs[:, 3:3 + H, pw >> 1, (pw & 1) * XP_C:(pw & 1) * XP_C + C] = np.transpose(x[:, :, :, w], (0, 2, 1))
    return xs


def pack_weights(w):
    co, C = w.shape[:2]
    rows = np.zeros((co, 7 * KROW))
    for kh in range(7):
        for kw in range(7):
            rows[:, kh * KROW + kw * XP_C:kh * KROW + kw * XP_C + C] = w[:, :, kh, kw]
    return rows


def conv1_s2d(x, w, pitch=64):
    B = x.shape[0]
    xs, wr = pack_input(x, pitch), pack_weights(w)
    out = np.zeros((B, 128, 128, w.shape[0]))
    for b in range(B):
        for oh in range(128):
            acc = np.zeros((128, w.shape[0]))
            for kh in range(7):
                box = np.zeros((LINES, 64))
                box[:, :pitch] = xs[b, 2 * oh + kh, :LINES, :]            # TMA box {pitch, 131}: the rest of each 128-byte line is never read
                for j in range(11):
                    tap, sub = j // 3, j % 3
                    a = box[tap:tap + 128, sub * 16:sub * 16 + 16]
                    assert sub * 16 + 16 <= 48 <= pitch
                    acc += a @ wr[:, kh * KROW + 16 * j:kh * KROW + 16 * j + 16].T
            out[b, oh] = acc
    return np.transpose(out, (0, 3, 1, 2))


def check_fused_pool(grid=148, B=3):
    """The POOL epilogue of conv1_s2d_kernel replayed per CTA: contiguous even-aligned ranges of conv rows with the odd row above
    recomputed for its carry; even lanes combine lanes -1 / +1 (pixel -1 = padding); an even row joins the running vertical maximum,
    an odd row completes pooled row (oh - 1) / 2 and becomes the row above of the next one -> must equal max_pool2d(3, 2, 1)."""
    rng = np.random.RandomState(1)
    y = np.maximum(rng.normal(0, 1, (B, 128, 128, 4)), 0)                      # relu(bn(conv1)) rows, NHWC
    n_rows = B * 128
    per = -(-n_rows // grid)
    per += per & 1
    out = np.full((B, 64, 64, 4), np.nan)
    stored = 0

    def hpool(v):                                                             # lanes 2j: max(v[2j-1], v[2j], v[2j+1])
        up = np.vstack([np.full((1, v.shape[1]), -np.inf), v[:-1]])
        dn = np.vstack([v[1:], np.full((1, v.shape[1]), -np.inf)])
        return np.maximum(np.maximum(up, v), dn)[0::2]
    for cta in range(grid):
        first = cta * per
        it1 = min(first + per, n_rows)
        it0 = first - (1 if (first & 127) != 0 and first < it1 else 0)
        vmax = np.full((64, 4), -np.inf)
        for item in range(it0, it1):
            b, oh = item >> 7, item & 127
            h = hpool(y[b, oh])
            if oh & 1:
                o = np.maximum(vmax, h)
                vmax = h
                if item >= first:
                    assert np.isnan(out[b, oh >> 1]).all()
                    out[b, oh >> 1] = o
                    stored += 1
            else:
                vmax = h if oh == 0 else np.maximum(vmax, h)
    ref = F.max_pool2d(torch.from_numpy(np.transpose(y, (0, 3, 1, 2))), 3, 2, 1).numpy()
    return stored == B * 64 and bool(np.array_equal(np.transpose(out, (0, 3, 1, 2)), ref))


def main():
    rng = np.random.RandomState(0)
    worst = 0.0
    for C, pitch in ((17, 64), (18, 64), (17, 48)):
        x = rng.normal(0, 1, (1, C, 256, 256))
        w = rng.normal(0, 1, (4, C, 7, 7))
        got = conv1_s2d(x, w, pitch)
        ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), stride=2, padding=3).numpy()
        err = float(np.abs(got - ref).max())
        worst = max(worst, err)
        print('C=%d pitch=%d: max abs error %.2e (output %s)' % (C, pitch, err, got.shape))
    a_now, w_now, a_new, w_new = 21 * 2 * 128 * 128, 21 * 2 * 64 * 128, 7 * 2 * LINES * 96, 21 * 128 * 128
    print('bytes per output row: A %d -> %d KB, W %d -> %d KB; TMA operations 84 -> 35'
          % (a_now // 1024, a_new // 1024, w_now // 1024, w_new // 1024))
    pool = check_fused_pool() and check_fused_pool(grid=148, B=64 // 8) and check_fused_pool(grid=5, B=1) and check_fused_pool(grid=148, B=1)
    print('fused max-pool epilogue (contiguous even-aligned row ranges + running vertical maximum) == max_pool2d(3, 2, 1):', pool)
    if worst > 1e-9 or not pool:
        print('MISMATCH')
        sys.exit(1)
    print('pair-layout conv1 == conv2d (max abs error %.2e)' % worst)


if __name__ == '__main__':
    main()
