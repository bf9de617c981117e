"""BASELINE config 5: LBS-only kernel sweep, achieved HBM GB/s against the measured roofline.

Algorithmic bytes per launch: 20,587,320 B of constants + 84,664 B per body.
Usage: python tools/bench_lbs.py [--batches 1 8 64 256 1024 4096] [--iters 50]   -> one JSON line per batch size
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'straps-3dhumanshapepose_b200'))
os.environ.setdefault('STRAPS_ASSETS_ROOT', os.path.join(REPO, 'tests', '_scratch', 'assets'))
import numpy as np
import torch
from straps_b200 import synthetic_assets

synthetic_assets.write_synthetic_assets(os.environ['STRAPS_ASSETS_ROOT'], seed=0)
import config
from models.smpl_official import SMPL
from utils.rigid_transform_utils import rot6d_to_rotmat

CONST_BYTES, BODY_BYTES = 20587320, 84664


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', type=int, nargs='+', default=[1, 8, 64, 256, 1024, 4096])
    ap.add_argument('--iters', type=int, default=50)
    ap.add_argument('--modes', nargs='+', default=['default'], help="STRAPS_LBS values to time: default, tc (tensor cores at every batch), simt")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(REPO, 'MEASURED_PEAKS.json')) else {}
    hbm = peaks.get('hbm_gbs', 6650.0)
    dev = 'cuda:0'
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # 256 MB > 126 MB L2
    for B, mode in [(b, m) for b in args.batches for m in args.modes]:
        if mode == 'default':
            os.environ.pop('STRAPS_LBS', None)
        else:
            os.environ['STRAPS_LBS'] = mode
        smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(dev)
        rng = np.random.RandomState(B)
        betas = torch.from_numpy(rng.normal(0, 1, (B, 10)).astype(np.float32)).to(dev)
        R = rot6d_to_rotmat(torch.from_numpy(rng.normal(0, 1, (B, 144)).astype(np.float32)).to(dev)).view(B, 24, 3, 3)
        go, bp = R[:, :1], R[:, 1:]
        with torch.no_grad():
            for _ in range(5):
                smpl(body_pose=bp, global_orient=go, betas=betas, pose2rot=False)
            cold, warm = [], []
            for it in range(args.iters):
                flush.zero_()                                   # evict L2 between timed launches
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                smpl(body_pose=bp, global_orient=go, betas=betas, pose2rot=False)
                e1.record()
                torch.cuda.synchronize()
                cold.append(e0.elapsed_time(e1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(args.iters):
                smpl(body_pose=bp, global_orient=go, betas=betas, pose2rot=False)
            e1.record()
            torch.cuda.synchronize()
            warm_ms = e0.elapsed_time(e1) / args.iters
        cold_ms = float(np.median(cold))
        bytes_ = CONST_BYTES + BODY_BYTES * B
        # The two figures above contain the host side of SMPL.forward (Python + ctypes + two launches, ~35 us): below B ~ 64 the GPU
        # waits for the host.  A CUDA graph of K forwards (with and without an L2 flush before each) minus a graph of the K flushes
        # alone gives the device time of the two kernels by themselves.
        graph = {}
        try:
            K = 20

            def timed_graph(with_fwd, with_flush):
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s), torch.no_grad():
                    with torch.cuda.graph(g, stream=s):
                        for _ in range(K):
                            if with_flush:
                                flush.zero_()
                            if with_fwd:
                                smpl(body_pose=bp, global_orient=go, betas=betas, pose2rot=False)
                torch.cuda.current_stream().wait_stream(s)
                g.replay()
                torch.cuda.synchronize()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(5):
                    g.replay()
                t1.record()
                torch.cuda.synchronize()
                return t0.elapsed_time(t1) / (5 * K)
            flush_ms = timed_graph(False, True)
            graph = {'ms_graph_cold_l2': timed_graph(True, True) - flush_ms, 'ms_graph_back_to_back': timed_graph(True, False)}
            graph['achieved_gbs_graph_cold'] = bytes_ / (graph['ms_graph_cold_l2'] * 1e-3) / 1e9
            graph['frac_graph_cold'] = graph['achieved_gbs_graph_cold'] / hbm
        except Exception as e:                       # capture is an extra: the eager figures stand on their own
            graph = {'graph_error': str(e)[:200]}
        print(json.dumps({'workload': 'SMPL LBS forward (LBS kernels + joints_kernel)', 'lbs_mode': mode, 'batch': B,
                          'ms_cold_l2': cold_ms, 'ms_back_to_back': warm_ms, 'bodies_per_s': B / (warm_ms * 1e-3),
                          'algorithmic_bytes': bytes_, 'achieved_gbs_cold': bytes_ / (cold_ms * 1e-3) / 1e9,
                          'achieved_gbs_back_to_back': bytes_ / (warm_ms * 1e-3) / 1e9, 'peak_gbs': hbm,
                          'frac_cold': bytes_ / (cold_ms * 1e-3) / 1e9 / hbm,
                          'fma_gflops': 2 * 8.56e6 * B / (warm_ms * 1e-3) / 1e9, **graph}))


if __name__ == '__main__':
    main()
