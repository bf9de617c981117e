"""Analytic model of conv_tc_kernel at B=64 against the measured launch list (no GPU needed).

Two measured constants explain every layer (profiles/r01_conv_experiments_s2.txt):
  * the TMA stream delivers ~56 B/clk into each SM in situ (conv1: 48 KB per 870 cycles with the MMAs switched off), and
  * an SS-mode tcgen05.mma M128 x N x K16 costs 48 / 64 / 128 cycles for N = 64 / 128 / 256 (profiles/r01_tc_probe.txt).
Per layer the model takes the slower of the two streams per tile plus a fixed per-tile epilogue exposure, times the number of
tile rounds on 148 SMs, and prints it beside the measured time.  It then prices the round-2 options of DESIGN.md 4.2 with the same
constants (halo tiles for the stride-1 3x3 layers), which is where the "~10 % of the step" estimate comes from.  The constant is bytes
INTO an SM: fetching half of each weight tile and multicasting it inside a 2-CTA cluster halves the L2 reads of W but not the bytes
each SM receives, and measured 1.6 % (a model that counted it as a saving would have predicted 20-25 % on layers 3-4).

    python tools/conv_model.py [--launches profiles/r01_launches_s2.txt]
"""
import argparse
import math
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMS, GHZ = 148, 1.965
TMA_B_PER_CLK = 56.0          # measured in situ, bytes per clock per SM
MMA_CYCLES = {64: 48, 128: 64, 256: 128}
EPILOGUE_EXPOSED = 1500       # cycles per tile not hidden behind the next tile's main loop (fitted on the short-K layers)
B = 64

# (name, cin, cout, ksize, stride, hout)  -- models/resnet.py:145-156,177-199 of the reference
LAYERS = [('conv1', None, 64, 7, 2, 128)]
cin = 64
for L, cout in enumerate((64, 128, 256, 512)):
    h = 64 >> L
    for blk in range(2):
        s = 2 if (blk == 0 and L > 0) else 1
        LAYERS.append(('layer%d.%d.conv1' % (L + 1, blk), cin, cout, 3, s, h))
        if blk == 0 and L > 0:
            LAYERS.append(('layer%d.%d.downsample' % (L + 1, blk), cin, cout, 1, 2, h))
        LAYERS.append(('layer%d.%d.conv2' % (L + 1, blk), cout, cout, 3, 1, h))
        cin = cout


def tile_model(cin, cout, ksize, stride, hout, halo=False):
    """-> (tiles, cycles per tile, bytes per tile, mma cycles per tile) of one convolution."""
    bn = 64 if cout == 64 else 128
    n_nt = cout // bn
    if cin is None:                                   # conv1: 7 filter rows x 3 chunks of 64 (168 real of 192), 11 of 12 K16 steps
        kblocks, ksteps = 21, 7 * 11
        a_bytes = kblocks * 2 * 128 * 128
    else:
        kblocks = ksize * ksize * cin // 64
        ksteps = kblocks * 4
        a_bytes = kblocks * 2 * 128 * 128
    positions = B * hout * hout
    if halo:                                          # padded raster, one halo tile per 64-channel chunk instead of one A tile per tap
        wp = hout + 2
        positions = B * wp * wp
        a_bytes = (cin // 64) * 2 * (128 + 2 * wp + 2) * 128
    w_bytes = kblocks * 2 * bn * 128
    tiles = math.ceil(positions / 128) * n_nt
    mma = ksteps * (MMA_CYCLES[2 * bn if 2 * bn <= 256 else bn] + MMA_CYCLES[bn])      # A_hi.[W_hi;W_lo] + A_lo.W_hi per K16
    tma = (a_bytes + w_bytes) / TMA_B_PER_CLK
    return tiles, max(mma, tma) + EPILOGUE_EXPOSED, a_bytes + w_bytes, mma


def layer_us(*args, **kw):
    tiles, cyc, _, _ = tile_model(*args, **kw)
    return math.ceil(tiles / SMS) * cyc / (GHZ * 1e3)


def port_model(cin, cout, ksize, stride, hout, variant='shipped'):
    """Second model: ONE shared-memory port of 128 B/clk per SM carries both the operand bytes the SS-mode MMAs read
    and the bytes TMA writes; cycles per tile = (read + written) / 128 + the exposed epilogue.  Per K16 step an MMA reads its A rows
    (128 x 32 B) and its B rows (N x 32 B).  -> (tiles, cycles per tile, KB read, KB written) for a kernel variant:
      shipped  conv_tc_kernel: wide A_hi.[W_hi;W_lo] (N = 2 bn) + narrow A_lo.W_hi (N = bn); stage = 2 A planes + 2 W planes
      pair     conv_tc2m_kernel: the same two MMAs with M = 256 over a CTA pair -- per SM the B rows halve, in reads and in writes
      halo     conv_halo_kernel: same reads; A written once per channel chunk as a (rows x (W+2))-line box
      s2d / s2d2  conv1_s2d_kernel: same reads; A = 7 boxes of 131 lines x 2 planes, W = 21 chunks of 16 KB (shared by two rows for s2d2)"""
    bn = 64 if cout == 64 else 128
    n_nt = cout // bn
    if cin is None:
        kblocks, ksteps = 21, 77
    else:
        kblocks = ksize * ksize * cin // 64
        ksteps = kblocks * 4
    positions = B * hout * hout
    b_rows = (2 * bn + bn) if variant != 'pair' else (bn + bn // 2)
    read = ksteps * (2 * 128 + b_rows) * 32
    a_w = kblocks * 2 * 128 * 128
    w_w = kblocks * 2 * bn * 128 if variant != 'pair' else kblocks * (bn + bn // 2) * 128
    if variant == 'halo':
        wp = hout + 2
        rh = -(-(128 + 2 * wp + 2) // wp) + 1
        positions = B * ((hout * wp + hout) // 128 + 1) * 128
        a_w = (cin // 64) * 2 * rh * wp * 128
    if variant in ('s2d', 's2d2'):
        a_w = 7 * 2 * 131 * 128
        w_w = 21 * 128 * 128 // (2 if variant == 's2d2' else 1)
    tiles = math.ceil(positions / 128) * n_nt
    return tiles, (read + a_w + w_w) / 128.0 + EPILOGUE_EXPOSED, read / 1024.0, (a_w + w_w) / 1024.0


def port_us(*args, **kw):
    tiles, cyc, _, _ = port_model(*args, **kw)
    return math.ceil(tiles / SMS) * cyc / (GHZ * 1e3)


def measured(path):
    out = []
    for line in open(path):
        m = re.search(r'conv_tc_kernel<.*?\s+([0-9.]+)\s+[0-9.]+%\s*$', line)
        if m:
            out.append(float(m.group(1)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--launches', default=os.path.join(REPO, 'profiles', 'r01_launches_s2.txt'))
    args = ap.parse_args()
    meas = measured(args.launches) if os.path.exists(args.launches) else []
    print('%-24s %7s %9s %9s %9s %9s | %9s' % ('layer', 'tiles', 'KB/tile', 'mma cyc', 'model us', 'ncu us', 'halo us'))
    tot = [0.0] * 3
    for i, (name, ci, co, k, s, h) in enumerate(LAYERS):
        tiles, cyc, nbytes, mma = tile_model(ci, co, k, s, h)
        us = layer_us(ci, co, k, s, h)
        halo_ok = ci is not None and k == 3 and s == 1 and h >= 32          # layers 1-2: border waste 6 % / 13 %
        us_h = layer_us(ci, co, k, s, h, halo=True) if halo_ok else us
        m = meas[i] if i < len(meas) else float('nan')
        print('%-24s %7d %9.0f %9d %9.1f %9.1f | %9.1f' % (name, tiles, nbytes / 1024, mma, us, m, us_h))
        for j, v in enumerate((us, m, us_h)):
            tot[j] += v
    print('%-24s %7s %9s %9s %9.1f %9.1f | %9.1f' % ('20 convolutions', '', '', '', *tot))
    print('model / measured = %.3f (the small stride-2 / 1x1 launches carry fixed costs the model ignores);  halo tiles on layers 1-2: '
          '-%.0f us of the step (model)' % (tot[0] / tot[1] if tot[1] else float('nan'), tot[0] - tot[2]))
    # ---- shared-memory port model: what each kernel behind a switch should buy if the port is the limit
    print()
    print('shared-memory port model: cycles per tile = (operand bytes the MMAs read + bytes TMA writes) / 128 B/clk + %d' % EPILOGUE_EXPOSED)
    print('%-24s %9s %9s %9s %9s | %9s %9s %9s' % ('layer', 'KB read', 'KB writ.', 'model us', 'ncu us', 'pair us', 'halo us', 's2d2 us'))
    tp = [0.0] * 5
    for i, (name, ci, co, k, s, h) in enumerate(LAYERS):
        _, _, rd, wr = port_model(ci, co, k, s, h)
        us = port_us(ci, co, k, s, h)
        halo_ok = ci is not None and k == 3 and s == 1 and h >= 32 and ci == co
        us_pair = port_us(ci, co, k, s, h, variant='pair')
        us_halo = port_us(ci, co, k, s, h, variant='halo') if halo_ok else us
        us_s2d = port_us(ci, co, k, s, h, variant='s2d2') if ci is None else us
        m = meas[i] if i < len(meas) else float('nan')
        print('%-24s %9.0f %9.0f %9.1f %9.1f | %9.1f %9.1f %9.1f' % (name, rd, wr, us, m, us_pair, us_halo, us_s2d))
        for j, v in enumerate((us, m, us_pair, us_halo, us_s2d)):
            tp[j] += v
    print('%-24s %9s %9s %9.1f %9.1f | %9.1f %9.1f %9.1f' % ('20 convolutions', '', '', *tp))
    print('port model / measured = %.3f.  Predictions: CTA pairs with the merged wide MMA -%.0f us, conv1 from the pair layout (two rows per '
          'item) -%.0f us,\nhalo tiles -%.0f us -- but the halo kernel MEASURED +20 .. +90 us (profiles/r01_halo_check.json), so for it '
          'something other than the port binds (its 3-stage weight ring: DESIGN.md 4.2 item 5).'
          % (tp[0] / tp[1] if tp[1] else float('nan'), tp[0] - tp[2], tp[0] - tp[4], tp[0] - tp[3]))


if __name__ == '__main__':
    main()
