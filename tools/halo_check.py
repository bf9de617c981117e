"""One-process hardware check of the convolution variants behind switches (conv_halo_kernel, EPW = 8, conv1_s2d_kernel) against the
shipped path (also the second CTA-pair kernel, conv_tc2m_kernel).

    python tools/halo_check.py [--out gpurun_out/halo_check.json] [--batch 8] [--time-batch 64] [--iters 20] [--only name,name]
    STRAPS_TC_S2D_PITCH=48 python tools/halo_check.py --only conv1_s2d,conv1_s2d2     (dense pair lines: a separate process)

The switches are read by the library once per encoder call (csrc/conv_tc.cu: tc_refresh_switches), so one process runs the shipped configuration and every
variant on the same weights and input.  The JSON is rewritten after every stage: a variant that hangs or faults still leaves the
stages before it (and its own name under "reached") on disk.  For every variant: the error of the encoder features and of each
layer1 / layer2 activation against the shipped path (max-abs / max-abs, the tolerance of tests/test_gpu_regressor.py), a per-pixel
error map of the first mismatching layer of images 0 and 1 (to see border / tile-boundary patterns), and CUDA-event timings.
The halo variants change the K order (chunk-major instead of tap-major), so they are NOT bit-identical: expected ~1e-6.
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, 'straps-3dhumanshapepose_b200')
ORACLE = os.path.join(REPO, 'oracle')
ASSETS = os.path.join(REPO, 'tests', '_scratch', 'assets')
os.environ.setdefault('STRAPS_ASSETS_ROOT', ASSETS)
for p in (ORACLE, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

VARIANTS = (('halo2', {'STRAPS_TC_HALO': '2'}), ('halo64', {'STRAPS_TC_HALO': '64'}), ('halo128', {'STRAPS_TC_HALO': '128'}), ('halo', {'STRAPS_TC_HALO': '1'}),
            ('halo_epw8', {'STRAPS_TC_HALO': '1,8'}), ('epw8', {'STRAPS_TC_EPI_WARPS': '8'}),
            ('conv1_s2d', {'STRAPS_TC_CONV1': 's2d'}), ('conv1_s2d2', {'STRAPS_TC_CONV1': 's2d2'}), ('conv1_s2dp', {'STRAPS_TC_CONV1': 's2dp'}),
            ('pair', {'STRAPS_TC_PAIR': 'all'}), ('pair_m128', {'STRAPS_TC_PAIR': 'm128'}), ('pair_m', {'STRAPS_TC_PAIR': 'm'}),
            ('pair_m+s2dp', {'STRAPS_TC_PAIR': 'm', 'STRAPS_TC_CONV1': 's2dp'}), ('pdl', {'STRAPS_TC_PDL': '1'}), ('tma2', {'STRAPS_TC_TMA2': '1'}),
            ('pdl+s2dp', {'STRAPS_TC_PDL': '1', 'STRAPS_TC_CONV1': 's2dp'}))
LAYERS = ['stem', 'pool'] + ['layer%d.%d%s' % (L, b, s) for L in (1, 2, 3, 4) for b in (0, 1) for s in ('.a', '')]
SWITCHES = ('STRAPS_TC_HALO', 'STRAPS_TC_EPI_WARPS', 'STRAPS_TC_CONV1', 'STRAPS_TC_PAIR', 'STRAPS_TC_PDL', 'STRAPS_TC_TMA2')


def set_env(env):
    for k in SWITCHES:
        os.environ.pop(k, None)
    os.environ.update(env)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(REPO, 'gpurun_out', 'halo_check.json'))
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--time-batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    res = {'reached': 'import'}

    def flush():
        with open(args.out + '.tmp', 'w') as f:
            json.dump(res, f, indent=1)
        os.replace(args.out + '.tmp', args.out)

    flush()
    import numpy as np
    import torch
    from straps_b200 import synthetic_assets, synthetic_inputs
    synthetic_assets.write_synthetic_assets(ASSETS, seed=0)
    import straps_oracle as O
    from models.regressor import SingleInputRegressor
    C = 17
    sd = O.make_regressor_state(C, seed=9)
    reg = SingleInputRegressor(C, 18, 3)
    reg.load_state_dict(sd)
    reg = reg.cuda().eval()
    enc = reg.image_encoder
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(args.batch, C, seed=1)).cuda()
    variants = [v for v in VARIANTS if not args.only or v[0] in args.only.split(',')]

    def run(env, xin):
        set_env(env)
        with torch.no_grad():
            feat = enc(xin).clone()
        acts = {n: enc._engine.read_activation(n, xin.shape[0]) for n in LAYERS} if xin.shape[0] == args.batch else {}
        torch.cuda.synchronize()
        return feat, acts

    res['reached'] = 'shipped'
    flush()
    f0, a0 = run({}, x)
    res['shipped'] = {'feat_abs_sum': float(f0.abs().sum())}
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    for name, env in variants:
        res['reached'] = name
        flush()
        try:
            f1, a1 = run(env, x)
        except Exception as e:                         # a launch / encode error: recorded, the next variant still runs
            res[name] = {'error': str(e)[:500]}
            flush()
            continue
        r = {'feat': rel(f1, f0), 'bit_identical': bool(torch.equal(f1, f0)), 'layers': {n: rel(a1[n], a0[n]) for n in LAYERS}}
        bad = [n for n in LAYERS if not r['layers'][n] < 1e-5]
        if bad:
            n = bad[0]
            err = (a1[n] - a0[n]).abs().amax(dim=1) / a0[n].abs().max()          # [B, H, W] per-pixel error
            np.save(args.out.replace('.json', '_%s_%s_errmap.npy' % (name, n)), err[:2].cpu().numpy())
            r['first_bad_layer'] = n
            r['bad_pixels_img0'] = int((err[0] > 1e-5).sum())
            r['bad_rows_img0'] = [int(i) for i in torch.nonzero((err[0] > 1e-5).any(dim=1)).flatten()[:80]]
            r['bad_cols_img0'] = [int(i) for i in torch.nonzero((err[0] > 1e-5).any(dim=0)).flatten()[:80]]
            cerr = (a1[n] - a0[n]).abs().amax(dim=(0, 2, 3)) / a0[n].abs().max()
            r['bad_channels'] = [int(i) for i in torch.nonzero(cerr > 1e-5).flatten()[:130]]
        res[name] = r
        flush()

    # timings: whole encoder, CUDA events on the current stream, inputs larger than nothing in particular (relative numbers only)
    res['reached'] = 'timing'
    flush()
    xt = torch.from_numpy(synthetic_inputs.make_proxy_batch(args.time_batch, C, seed=2)).cuda()
    res['timing_ms'] = {}
    res['feat_at_time_batch'] = {}
    ft0 = None
    for name, env in [('shipped', {})] + [v for v in variants if 'error' not in res.get(v[0], {})]:
        res['reached'] = 'timing:' + name
        flush()
        set_env(env)
        with torch.no_grad():
            for _ in range(3):
                ft = enc(xt).clone()
            torch.cuda.synchronize()
            if ft0 is None:
                ft0 = ft
            # the comparison above runs at a small batch; this one is the bench's batch (more tiles per SM, every ring wraps many times)
            res['feat_at_time_batch'][name] = {'rel': rel(ft, ft0), 'bit_identical': bool(torch.equal(ft, ft0))}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.iters):
                enc(xt)
            e1.record()
            torch.cuda.synchronize()
        res['timing_ms'][name] = e0.elapsed_time(e1) / args.iters
        flush()
    res['reached'] = 'done'
    flush()
    print(json.dumps(res))


if __name__ == '__main__':
    main()
