#!/usr/bin/env python
"""bench.py -- bodies/sec of the STRAPS hot path (ResNet-18 encoder -> 3x IEF -> rot6d -> SMPL forward).

Workload (BASELINE.json configs[1]): B=64 synthetic 256x256x17 proxy representations per GPU, inference,
random-init weights of the reference architecture, seeded synthetic SMPL-shaped assets (the real SMPL files are
licence gated).  One "step" = one pass of the whole hot path over one batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--channels 17] [--conv-mode f16x3_tc]
    torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, weak scaling, no collective)
    python bench.py --impl reference ...                       (the reference's CPU PyTorch path, host cores)

Prints ONE JSON line (rank 0).  Keys follow the driver contract:
  value      bodies/s with the inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e        bodies/s through the public drop-in API with HOST (pinned) input: H2D of the batch + D2H of the results
             inside the timed region
  roofline   tensor-pipe roofline of the convolution stack (algorithmic FLOPs / measured time / measured bf16 peak)
  cpu_baseline  the oracle (CPU restatement of the reference path) timed on this box's host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, 'straps-3dhumanshapepose_b200')
ORACLE = os.path.join(REPO, 'oracle')
ASSETS = os.path.join(REPO, 'tests', '_scratch', 'assets')
os.environ.setdefault('STRAPS_ASSETS_ROOT', ASSETS)
if PKG not in sys.path:
    sys.path.insert(0, PKG)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = 'bodies/sec (SMPL fwd+regress, B=64, 256x256x17)'
# algorithmic FLOPs of encoder + IEF per body (hook counted on the reference modules)
MFLOP_PER_BODY = {17: 6180.1, 18: 6283.0}


def algorithmic_mflop(c):
    return MFLOP_PER_BODY.get(c, 4429.1 + 2 * 64 * 128 * 128 * 49 * c / 1e6 + 4.11)


def conv_traffic(batch, channels):
    """DRAM bytes per conv_tc_kernel launch from the committed `ncu --set full` capture of this command (B=64, C=17 only)."""
    p = os.path.join(REPO, 'profiles', 'r01_conv_traffic.json')
    if batch == 64 and channels == 17 and os.path.exists(p):
        return json.load(open(p))['dram_bytes_per_launch']
    return None


def measured_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1590.0)), d.get('hbm_gbs', 6650.0), 'measured'
    return 1590.0, 6650.0, 'fallback'


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference path), all host threads
# --------------------------------------------------------------------------------------------------
def cpu_path(batch, channels, reps, warm):
    if ORACLE not in sys.path:
        sys.path.insert(0, ORACLE)
    import straps_oracle as O
    from straps_b200 import synthetic_assets, synthetic_inputs
    synthetic_assets.write_synthetic_assets(ASSETS, seed=0)
    sd = O.make_regressor_state(channels, seed=1)
    add = os.path.join(ASSETS, 'additional')
    smpl = O.SmplOracle(add, batch_size=batch)
    init = O.load_initial_params(os.path.join(add, 'neutral_smpl_mean_params_6dpose.npz'))
    x = torch.from_numpy(synthetic_inputs.make_proxy_batch(batch, channels, seed=3))
    def one_pass():
        t0 = time.perf_counter()
        with torch.no_grad():
            O.regress_and_pose(x, sd, init, smpl)
        return time.perf_counter() - t0
    # the reference arm gets its best host configuration: probe a few intra-op thread counts, keep the fastest
    ncpu = os.cpu_count() or 1
    best, cores = None, ncpu
    for n in sorted({ncpu, max(1, ncpu // 2), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
        torch.set_num_threads(n)
        one_pass()
        t = one_pass()
        if best is None or t < best:
            best, cores = t, n
    torch.set_num_threads(cores)
    times = [one_pass() for _ in range(warm + reps)][warm:]
    return batch / float(np.median(times)), float(np.median(times)), cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = min(args.batch, args.cpu_sample)
    bps, sec, cores = cpu_path(sample, args.channels, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {'impl': 'reference', 'metric': METRIC, 'value': bps, 'unit': 'bodies/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'encoder+3xIEF+rot6d+SMPL fwd, inference, 256x256x%d' % args.channels,
                       'batch_per_step': sample, 'l2': 'n/a (CPU)'},
            'cpu_baseline': {'value': bps, 'unit': 'bodies/s', 'cores': cores, 'kind': 'port',
                             'sample': 'median of %d steps of one B=%d batch (oracle = CPU PyTorch restatement of the '
                                       'reference path, all host threads)' % (max(1, args.steps), sample)},
            'e2e': {'value': bps, 'unit': 'bodies/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'torch': torch.__version__}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from straps_b200 import synthetic_assets, synthetic_inputs, _lib
    if rank == 0:
        synthetic_assets.write_synthetic_assets(ASSETS, seed=0)
    if world > 1:
        dist.barrier()
    import config  # noqa: F401
    from models.regressor import SingleInputRegressor
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat

    B, C = args.batch, args.channels
    torch.manual_seed(0)   # identical replica weights on every rank
    reg = SingleInputRegressor(C, 18, 3, conv_mode=args.conv_mode).to(dev).eval()
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(dev)
    x_host = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=100 + rank)).pin_memory()
    x_dev = x_host.to(dev)

    def hot_path(x):
        cam, pose, shape = reg(x)
        R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
        return cam, out.vertices, out.joints

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            hot_path(x_dev)
        # ---- value: inputs resident in HBM; the 285 MB input + 1 GB of activations exceed the 126 MB L2 ----
        sampler = ClockSampler(local)
        barrier()
        if rank == 0:
            sampler.start()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        enc_ms = 0.0
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        e0.record()
        for i in range(args.steps):
            hot_path(x_dev)
        e1.record()
        barrier()
        launches = _lib.launch_count() - n0
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None
        # ---- roofline of the convolution stack: time the encoder alone, same stream, CUDA events ----
        for i in range(args.steps):
            ev[i][0].record()
            reg.image_encoder(x_dev)
            ev[i][1].record()
        barrier()
        enc_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        # ---- e2e: pinned host input -> H2D -> hot path -> D2H of every result, every step, through the public API.
        # The copy of step i+1 is issued on a second stream so that it overlaps the kernels of step i (two staging buffers),
        # the way a caller of a device-tensor API pipelines its input; every byte still moves inside the timed region.
        out_host = [torch.empty((B, 3)).pin_memory(), torch.empty((B, 6890, 3)).pin_memory(), torch.empty((B, 90, 3)).pin_memory()]
        stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
        copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        for ev_ in free:
            ev_.record(cur)

        def e2e_step(i):
            sidx = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[sidx])
                stage[sidx].copy_(x_host, non_blocking=True)
                ready[sidx].record(copy_stream)
            cur.wait_event(ready[sidx])
            res = hot_path(stage[sidx])
            free[sidx].record(cur)
            for h, d in zip(out_host, res):
                h.copy_(d, non_blocking=True)
        for i in range(4):
            e2e_step(i)
        barrier()
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))

        # ---- extra: the proxy representation is synthesised ON THE DEVICE from what a caller really has on
        # the host -- part-segmentation labels and 2-D joints (reference train/...:178-182) -- so only 16.8 MB + 8.7 KB cross PCIe.
        ms_kp = None
        if C in (17, 18):
            from utils.label_conversions import convert_2Djoints_to_gaussian_heatmaps_torch, convert_multiclass_to_binary_labels_torch
            rng = np.random.RandomState(7 + rank)
            seg_host = torch.from_numpy((x_host[:, 0].numpy() * rng.randint(1, 7, (B, 1, 1))).astype(np.float32)).pin_memory()
            j2d_host = torch.from_numpy(rng.uniform(8, 247, (B, C - 1, 2)).astype(np.float32)).pin_memory()
            seg_dev, j2d_dev = torch.empty_like(seg_host, device=dev), torch.empty_like(j2d_host, device=dev)

            def kp_step():
                seg_dev.copy_(seg_host, non_blocking=True)
                j2d_dev.copy_(j2d_host, non_blocking=True)
                x_in = torch.cat([convert_multiclass_to_binary_labels_torch(seg_dev).unsqueeze(1),
                                  convert_2Djoints_to_gaussian_heatmaps_torch(j2d_dev, 256)], dim=1)
                res = hot_path(x_in)
                for h, d in zip(out_host, res):
                    h.copy_(d, non_blocking=True)
            for _ in range(3):
                kp_step()
            barrier()
            e0.record()
            for _ in range(args.steps):
                kp_step()
            e1.record()
            barrier()
            ms_kp = max_over_ranks(e0.elapsed_time(e1))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e = world * B / (ms_e2e / args.steps * 1e-3)
    peak_tf, peak_hbm, peak_kind = measured_peaks()
    flops = algorithmic_mflop(C) * 1e6 * B
    achieved_tf = flops / (enc_ms * 1e-3) / 1e12
    cpu = None
    if world == 1 or True:
        sample = min(B, args.cpu_sample)
        bps, sec, cores = cpu_path(sample, C, args.cpu_reps, 1)
        cpu = {'value': bps, 'unit': 'bodies/s', 'cores': cores, 'kind': 'port',
               'sample': 'median of %d passes over one B=%d batch of the same workload (oracle = CPU PyTorch '
                         'restatement of the reference path, %d threads)' % (args.cpu_reps, sample, cores)}
    line = {'metric': METRIC, 'value': value, 'unit': 'bodies/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'fp16x3 split (fp32-equivalent) convs, fp32 elsewhere' if args.conv_mode == 'f16x3_tc' else 'f32',
            'data': 'synthetic',
            'config': {'workload': 'encoder+3xIEF+rot6d+SMPL fwd, inference, B=%d/GPU, 256x256x%d (BASELINE configs[1])' % (B, C),
                       'global_batch': world * B, 'conv_mode': args.conv_mode, 'parallelism': 'dp%d replicas, no collective' % world,
                       'l2': 'inputs+activations per step (>1 GB) exceed the 126 MB L2; no explicit flush'},
            'e2e': {'value': e2e, 'unit': 'bodies/s', 'h2d_bytes_per_step': int(x_host.numel() * 4),
                    'd2h_bytes_per_step': int(sum(h.numel() for h in out_host) * 4)},
            'e2e_from_keypoints': None if ms_kp is None else {
                'value': world * B / (ms_kp / args.steps * 1e-3), 'unit': 'bodies/s',
                'h2d_bytes_per_step': int(B * 256 * 256 * 4 + B * (C - 1) * 2 * 4), 'd2h_bytes_per_step': int(sum(h.numel() for h in out_host) * 4),
                'note': 'input synthesised on device from host segmentation labels + 2-D joints (utils/label_conversions drop-in, SURVEY 8f N1)'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': achieved_tf / peak_tf, 'traffic': conv_traffic(B, C), 'peak_kind': peak_kind + ' (sustained cuBLAS bf16)',
                         'kernel': 'conv_tc_kernel (20 launches/step) + input pack/pools = encoder', 'encoder_ms': enc_ms,
                         'algorithmic_gflop_per_step': flops / 1e9,
                         'traffic_unit': 'DRAM bytes per conv_tc_kernel launch (dram__bytes_read+write, mean of the 20 launches of a step)',
                         'note': 'algorithmic FLOPs (1 pass); the f16x3 mode issues 3 MMA passes'},
            'cpu_baseline': cpu, 'clocks': clocks, 'torch': torch.__version__}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--channels', type=int, default=17)
    ap.add_argument('--conv-mode', default=os.environ.get('STRAPS_CONV_MODE', 'f16x3_tc'))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-sample', type=int, default=64, help='bodies per CPU-baseline pass')
    ap.add_argument('--cpu-reps', type=int, default=3)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
