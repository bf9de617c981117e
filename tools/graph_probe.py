"""How much of a step is launch gaps?  Times the hot path (BASELINE configs[1]) launched eagerly through the drop-in API
against a CUDA-graph replay of the very same calls (torch.cuda.graph capture on a side stream).

Usage: python tools/graph_probe.py [--batch 64] [--steps 50]      -> one JSON line
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'straps-3dhumanshapepose_b200'))
os.environ.setdefault('STRAPS_ASSETS_ROOT', os.path.join(REPO, 'tests', '_scratch', 'assets'))
import torch
from straps_b200 import synthetic_assets, synthetic_inputs, _lib

synthetic_assets.write_synthetic_assets(os.environ['STRAPS_ASSETS_ROOT'], seed=0)
import config
from models.regressor import SingleInputRegressor
from models.smpl_official import SMPL
from utils.rigid_transform_utils import rot6d_to_rotmat


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--steps', type=int, default=50)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    B, C = args.batch, 17
    torch.manual_seed(0)
    reg = SingleInputRegressor(C, 18, 3).to(dev).eval()
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(dev)
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=100)).to(dev)

    def hot_path(inp):
        cam, pose, shape = reg(inp)
        R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
        return cam, out.vertices, out.joints

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    with torch.no_grad():
        for _ in range(5):
            ref = hot_path(x)
        eager_ms = timed(lambda: hot_path(x))
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            hot_path(x)
        torch.cuda.current_stream().wait_stream(s)
        n0 = _lib.launch_count()
        with torch.cuda.graph(g):
            out = hot_path(x)
        captured = _lib.launch_count() - n0
        g.replay()
        torch.cuda.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(ref, out))
        graph_ms = timed(g.replay)
    print(json.dumps({'workload': 'encoder+3xIEF+rot6d+SMPL fwd, B=%d, eager launches vs CUDA-graph replay' % B, 'eager_ms': eager_ms,
                      'graph_ms': graph_ms, 'launch_gap_share': 1 - graph_ms / eager_ms, 'library_kernels_captured': captured,
                      'graph_outputs_bit_identical': same}))


if __name__ == '__main__':
    main()
