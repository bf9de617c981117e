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
  train      BASELINE config 3 (N=1) / config 4 (N>1): the full training step -- forward, multi-task loss, backward, ONE NCCL all-reduce
             of the flat gradient bucket, fused Adam -- in the shipped tensor-core mode, device-timed, max over ranks
  lbs_sweep  BASELINE config 5 (N=1 only): the LBS kernels alone at B = 1 .. 4096, achieved GB/s of the algorithmic bytes vs measured HBM peak
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
    """DRAM bytes per convolution-kernel launch from the committed `ncu --set full` capture of this command (B=64, C=17 only);
    regenerated whenever the convolution path changes (profiles/r02_conv_traffic.json, else null)."""
    d = conv_traffic_r02()
    if batch == 64 and channels == 17 and d is not None:
        return d['dram_bytes_per_launch']
    return None


def measured_peaks():
    """(bf16 TFLOP/s, HBM GB/s, provenance).  The timed region is tens of milliseconds at maximum clocks, i.e. burst conditions, so the
    tensor roofline is taken against the BURST cuBLAS figure of MEASURED_PEAKS.json (round-1 review), not the sustained one."""
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops', 1590.0), d.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json: burst cuBLAS bf16, copy bandwidth)'
    return 1590.0, 6650.0, 'fallback (B200_PROFILING.md)'


def conv_traffic_r02():
    p = os.path.join(REPO, 'profiles', 'r02_conv_traffic.json')
    return json.load(open(p)) if os.path.exists(p) else None


def workload_config(batch, channels, n_gpus, conv_mode):
    """The same dict in both arms (the driver compares them)."""
    return {'workload': 'encoder+3xIEF+rot6d+SMPL fwd, inference, B=%d per GPU, 256x256x%d (BASELINE configs[1])' % (batch, channels),
            'global_batch': n_gpus * batch, 'batch_per_gpu': batch, 'channels': channels, 'parallelism': 'dp%d replicas, no collective' % n_gpus,
            'l2': 'inputs+activations per step (>1 GB) exceed the 126 MB L2; no explicit flush'}


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore the pinned host buffers it allocates next: first touch) to the NUMA node the GPU hangs off.
    Round 1's e2e curve collapsed at 4-8 GPUs because every rank's 285 MB/step H2D copy read host memory of whichever node the
    allocation happened to land on.  Best effort: returns what was done for the JSON line."""
    info = {'bound': False}
    try:
        prop = torch.cuda.get_device_properties(index)
        bus = '%04x:%02x:%02x.0' % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
        info.update(pci=bus, node=node)
        cpus = set()
        if node >= 0:
            for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        else:
            # sysfs hides the topology (containers, some VMs): ask NVML for the GPU's ideal CPU set instead
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
            info['source'] = 'nvml'
            if len(cpus) >= (os.cpu_count() or 1):
                info['note'] = 'NVML reports no CPU locality for this GPU (single NUMA domain)'
                return info
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus))
        if node < 0:
            return info
        try:                                   # also prefer the node for page allocation (set_mempolicy(MPOL_PREFERRED)); x86-64 syscall 238
            import ctypes
            mask = ctypes.c_ulong(1 << node)
            rc = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
            info['mempolicy'] = 'preferred' if rc == 0 else 'errno %d' % ctypes.get_errno()
        except Exception as e:                 # noqa: BLE001
            info['mempolicy'] = 'unavailable (%s)' % type(e).__name__
    except Exception as e:                     # noqa: BLE001 -- sysfs layout differs, containers hide it: the run goes on unbound
        info['error'] = '%s: %s' % (type(e).__name__, str(e)[:80])
    return info


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
            'config': workload_config(args.batch, args.channels, args.gpus, args.conv_mode),
            'cpu_baseline': {'value': bps, 'unit': 'bodies/s', 'cores': cores, 'kind': 'port',
                             'sample': 'median of %d steps of one B=%d batch (oracle = CPU PyTorch restatement of the '
                                       'reference path, all host threads)' % (max(1, args.steps), sample)},
            'e2e': {'value': bps, 'unit': 'bodies/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'torch': torch.__version__}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
LBS_CONST_BYTES, LBS_BODY_BYTES = 20587320, 84664      # algorithmic bytes of one LBS launch = constants + per body


def lbs_sweep(dev, batches, peak_hbm):
    """BASELINE config 5: the LBS kernels alone (SMPL.forward with rotation matrices), device time from CUDA-graph replays with the L2
    flushed before every launch (a graph of the flushes alone is subtracted), achieved GB/s of the algorithmic bytes."""
    import config
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)          # 256 MB > 126 MB L2
    rows = []
    for B in batches:
        smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(dev)
        rng = np.random.RandomState(B)
        betas = torch.from_numpy(rng.normal(0, 1, (B, 10)).astype(np.float32)).to(dev)
        R = rot6d_to_rotmat(torch.from_numpy(rng.normal(0, 1, (B, 144)).astype(np.float32)).to(dev)).view(B, 24, 3, 3)
        go, bp = R[:, :1], R[:, 1:]
        K = 10

        def timed_graph(with_fwd, with_flush):
            g = torch.cuda.CUDAGraph()
            st = torch.cuda.Stream(device=dev)
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st), torch.no_grad():
                for _ in range(2):
                    smpl(body_pose=bp, global_orient=go, betas=betas, pose2rot=False)
                with torch.cuda.graph(g, stream=st):
                    for _ in range(K):
                        if with_flush:
                            flush.zero_()
                        if with_fwd:
                            smpl(body_pose=bp, global_orient=go, betas=betas, pose2rot=False)
            torch.cuda.current_stream(dev).wait_stream(st)
            g.replay()
            torch.cuda.synchronize(dev)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(3):
                g.replay()
            t1.record()
            torch.cuda.synchronize(dev)
            return t0.elapsed_time(t1) / (3 * K)
        flush_ms = timed_graph(False, True)
        cold_ms = max(timed_graph(True, True) - flush_ms, 1e-6)
        warm_ms = timed_graph(True, False)
        nbytes = LBS_CONST_BYTES + LBS_BODY_BYTES * B
        rows.append({'batch': B, 'us_cold_l2': cold_ms * 1e3, 'us_back_to_back': warm_ms * 1e3, 'bodies_per_s': B / (warm_ms * 1e-3),
                     'algorithmic_bytes': nbytes, 'achieved_gbs': nbytes / (cold_ms * 1e-3) / 1e9,
                     'frac': nbytes / (cold_ms * 1e-3) / 1e9 / peak_hbm})
        del smpl
    return {'config': 'BASELINE configs[4]: SMPL LBS-only (lbs kernels + joints kernel), rotation-matrix input, L2 flushed between launches',
            'bound': 'hbm', 'peak': peak_hbm, 'unit': 'GB/s', 'rows': rows}


def train_bench(dev, rank, world, B, C, conv_mode, steps, x_dev, barrier, max_over_ranks, no_graph=False):
    """BASELINE config 3 (world 1) / config 4 (world > 1): reference train/train_synthetic_otf_rendering.py:186-233 without the renderer --
    regressor.train() forward, rot6d, SMPL, projection, five-term multi-task loss, backward, ONE all-reduce (NCCL) of the flat fp32
    gradient bucket, fused Adam (run_train.py:200-201) -- on device-resident synthetic inputs and targets."""
    import config
    from models.regressor import SingleInputRegressor
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat
    from utils.cam_utils import orthographic_project_torch
    from utils.joints2d_utils import check_joints2d_visibility_torch
    from losses.multi_task_loss import HomoscedasticUncertaintyWeightedMultiTaskLoss as Loss
    from straps_b200.parallel import DataParallelAdam
    from straps_b200 import _lib
    from straps_b200.ops import select_joints
    tasks = ['verts', 'joints2D', 'joints3D', 'pose_params', 'shape_params']
    weights = {'verts': 1.0, 'joints2D': 0.1, 'pose_params': 0.1, 'shape_params': 0.1, 'joints3D': 1.0}      # run_train.py:53-54
    torch.manual_seed(1)
    reg = SingleInputRegressor(C, 18, 3, conv_mode=conv_mode).to(dev).train()
    crit = Loss(tasks, init_loss_weights=weights).to(dev)
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(dev)
    opt = DataParallelAdam(list(reg.parameters()) + list(crit.parameters()), lr=1e-4)
    # STRAPS_DP_OVERLAP=1: start the all-reduce of layer4 + IEF inside the backward pass (DataParallelAdam.enable_overlap).  Off by
    # default: measured on 8 x B200 it changes nothing (8.04 vs 8.05 ms, profiles/r02_train_8gpu_overlap.jsonl)
    if os.environ.get('STRAPS_DP_OVERLAP', '0') != '0':
        opt.enable_overlap(reg)
    rng = np.random.RandomState(500 + rank)
    with torch.no_grad():                                   # target side (train/...:121-145), once: seeded random pose / shape
        t_betas = torch.from_numpy(rng.normal(0, 1, (B, 10)).astype(np.float32)).to(dev)
        t_R = rot6d_to_rotmat(torch.from_numpy(rng.normal(0, 1, (B, 144)).astype(np.float32)).to(dev)).view(B, 24, 3, 3)
        t_out = smpl(body_pose=t_R[:, 1:], global_orient=t_R[:, :1], betas=t_betas, pose2rot=False)
        t_j2d = torch.from_numpy(rng.uniform(-20, 276, (B, 17, 2)).astype(np.float32)).to(dev)
        labels = {'verts': t_out.vertices, 'joints2D': t_j2d,
                  'joints3D': t_out.joints[:, config.ALL_JOINTS_TO_H36M_MAP, :][:, config.H36M_TO_J14, :].contiguous(),
                  'shape_params': t_betas, 'pose_params_rot_matrices': t_R,
                  'vis': check_joints2d_visibility_torch(t_j2d, config.REGRESSOR_IMG_WH)}

    def step():
        opt.zero_grad()
        cam, pose, shape = reg(x_dev)
        R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
        outs = {'verts': out.vertices, 'joints2D': orthographic_project_torch(select_joints(out.joints, config.ALL_JOINTS_TO_COCO_MAP), cam),
                'joints3D': select_joints(out.joints, config.ALL_JOINTS_TO_H36M_MAP, config.H36M_TO_J14), 'shape_params': shape,
                'pose_params_rot_matrices': R}
        loss, _ = crit(labels, outs)
        loss.backward()
        opt.step()                                          # all-reduce(sum) of the bucket + Adam with 1/world folded in
        return loss.detach()
    def timed(fn):
        for _ in range(3):
            out = fn()
        barrier()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, (_lib.launch_count() - n0) / steps, out
    ms_eager, launches, loss = timed(step)
    # the same step captured as CUDA graph(s) and replayed (straps_b200.graphs.GraphedTrainStep; the all-reduce stays outside the capture)
    ms_graph, graph_note = None, None
    try:
        if no_graph:
            raise RuntimeError('skipped (--no-train-graph)')
        from straps_b200.graphs import GraphedTrainStep
        gstep = GraphedTrainStep(step, opt)
        ms_graph, _, loss = timed(gstep)
    except Exception as e:                                  # noqa: BLE001 -- the eager figure stands on its own
        graph_note = '%s: %s' % (type(e).__name__, str(e)[:300])
        if world > 1 and not no_graph:
            raise                                           # a half-captured rank would dead-lock the others at the next collective
    ms = ms_graph if ms_graph is not None else ms_eager
    # identical replicas after the updates: every rank's parameter bucket must hash the same
    same = None
    if world > 1:
        import torch.distributed as dist
        h = opt.bucket.params.double().sum().reshape(1)
        hs = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        same = bool(all(torch.equal(hs[0], t) for t in hs))
    flops = 3 * algorithmic_mflop(C) * 1e6 * B             # forward + data gradient + weight gradient
    peak_tf = measured_peaks()[0]
    out = {'config': 'BASELINE configs[%d]: training step (encoder+IEF+rot6d+SMPL+projection+multi-task loss, no renderer), B=%d per GPU, '
                     '%d GPU(s), conv mode %s, %s, fused Adam'
                     % (3 if world > 1 else 2, B, world, conv_mode,
                        ('NCCL all-reduce of %d fp32 gradients per step%s' % (opt.bucket.numel, ', layer4 + IEF part (%d) started inside the backward pass on a side stream' % (opt._early_range[1] - opt._early_range[0]) if opt._early_range else '')) if world > 1 else 'no collective (1 GPU)'),
           'ms_per_step': ms, 'value': world * B / (ms * 1e-3), 'unit': 'bodies/s', 'steps': steps, 'global_batch': world * B,
           'mode': 'CUDA-graph replay of the whole step (GraphedTrainStep)' if ms_graph is not None else 'eager PyTorch loop',
           'ms_per_step_eager': ms_eager, 'ms_per_step_graphed': ms_graph, 'graph_note': graph_note,
           'library_launches_per_step': launches, 'final_loss': float(loss.detach()), 'optimiser_steps': opt.step_count,
           'allreduce_elements': opt.bucket.numel if world > 1 else 0, 'replicas_identical_after_update': same,
           'algorithmic_tflops_per_gpu': flops / (ms * 1e-3) / 1e12, 'frac_of_bf16_peak': flops / (ms * 1e-3) / 1e12 / peak_tf}
    del reg, crit, smpl, opt
    torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa_node(local) if not args.no_numa_bind else {'bound': False, 'note': '--no-numa-bind'}
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

    if args.train_only:             # development shortcut (multi-GPU A/B runs of the training arm): not the driver's line
        barrier0 = (lambda: (dist.barrier() if world > 1 else None, torch.cuda.synchronize()))

        def mx(ms):
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())
            return ms
        train = train_bench(dev, rank, world, B, C, args.conv_mode, args.train_steps, x_dev, barrier0, mx, no_graph=args.no_train_graph)
        if rank == 0:
            print(json.dumps({'train': train, 'n_gpus': world}))
        if world > 1:
            dist.destroy_process_group()
        return

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

    def min_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return float(t.item())
        return v

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
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        e0.record()
        for i in range(args.steps):
            hot_path(x_dev)
        e1.record()
        barrier()
        launches = _lib.launch_count() - n0
        ms_total_eager = max_over_ranks(e0.elapsed_time(e1))
        # ---- the same step as a CUDA-graph replay (straps_b200.graphs.GraphedCallable, the repo's public capture API): the ~28 launches
        # of a step leave ~2 % of host-induced gaps in the eager loop.  The replay reads the graph's own input buffer (inputs resident).
        ms_total, enc_graph, value_mode = ms_total_eager, None, 'eager PyTorch loop'
        if not args.no_graph:
            try:
                from straps_b200.graphs import GraphedCallable
                g_hot = GraphedCallable(hot_path, x_dev)
                xs = g_hot.static_inputs[0]
                for _ in range(3):
                    g_hot(xs)
                barrier()
                e0.record()
                for i in range(args.steps):
                    g_hot(xs)
                e1.record()
                barrier()
                ms_total = max_over_ranks(e0.elapsed_time(e1))
                value_mode = 'CUDA-graph replay of the step (GraphedCallable)'
                enc_graph = GraphedCallable(lambda t: reg.image_encoder(t), x_dev)
            except Exception as e:                                  # noqa: BLE001 -- the eager figure stands on its own
                value_mode = 'eager PyTorch loop (graph capture failed: %s)' % (str(e)[:120],)
                ms_total, enc_graph = ms_total_eager, None
        clocks = sampler.stop() if rank == 0 else None
        # ---- roofline of the convolution stack: time the encoder alone, same stream, CUDA events ----
        enc_fn, enc_in = (enc_graph, enc_graph.static_inputs[0]) if enc_graph is not None else (reg.image_encoder, x_dev)
        for i in range(3):
            enc_fn(enc_in)
        for i in range(args.steps):
            ev[i][0].record()
            enc_fn(enc_in)
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

        d2h_stream = torch.cuda.Stream(device=dev)
        computed = torch.cuda.Event()

        def read_back(res):
            # device -> host on a third stream: the 5.4 MB of results leave while the next step's kernels run (on the compute
            # stream the copy engine would hold up the following launch for ~0.1 ms per step)
            computed.record(cur)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(computed)
                for h, d in zip(out_host, res):
                    h.copy_(d, non_blocking=True)
                    d.record_stream(d2h_stream)

        def e2e_step(i):
            sidx = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[sidx])
                stage[sidx].copy_(x_host, non_blocking=True)
                ready[sidx].record(copy_stream)
            cur.wait_event(ready[sidx])
            res = hot_path(stage[sidx])
            free[sidx].record(cur)
            read_back(res)
        for i in range(4):
            e2e_step(i)
        barrier()
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        cur.wait_stream(d2h_stream)                          # the last step's results are on the host before the clock stops
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        # ---- the host->device copy alone, all ranks at once: the ceiling of the fp32-tensor e2e (285 MB per step per rank) ----
        barrier()
        e0.record()
        for i in range(10):
            stage[i % 2].copy_(x_host, non_blocking=True)
        e1.record()
        barrier()
        h2d_gbs = x_host.numel() * 4 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        h2d_min, h2d_max = min_over_ranks(h2d_gbs), -min_over_ranks(-h2d_gbs)
        del stage

        # ---- extra: the proxy representation is synthesised ON THE DEVICE from what a caller really has on
        # the host -- part-segmentation labels and 2-D joints (reference train/...:178-182) -- so only 16.8 MB + 8.7 KB cross PCIe,
        # and since round 2 it is generated INSIDE the stem's input pack (SingleInputRegressor.forward_from_labels: no fp32 tensor).
        ms_kp = None
        if C in (17, 18) and args.conv_mode == 'f16x3_tc':
            rng = np.random.RandomState(7 + rank)
            seg_host = torch.from_numpy((x_host[:, 0].numpy() * rng.randint(1, 7, (B, 1, 1))).astype(np.float32)).pin_memory()
            j2d_host = torch.from_numpy(rng.uniform(8, 247, (B, C - 1, 2)).astype(np.float32)).pin_memory()
            seg_dev = [torch.empty_like(seg_host, device=dev) for _ in range(2)]
            j2d_dev = [torch.empty_like(j2d_host, device=dev) for _ in range(2)]
            for ev_ in free:
                ev_.record(cur)

            def kp_step(i):
                sidx = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[sidx])
                    seg_dev[sidx].copy_(seg_host, non_blocking=True)
                    j2d_dev[sidx].copy_(j2d_host, non_blocking=True)
                    ready[sidx].record(copy_stream)
                cur.wait_event(ready[sidx])
                cam, pose, shape = reg.forward_from_labels(seg_dev[sidx], j2d_dev[sidx])
                free[sidx].record(cur)
                R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
                out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
                read_back((cam, out.vertices, out.joints))
            for i in range(4):
                kp_step(i)
            barrier()
            e0.record()
            for i in range(args.steps):
                kp_step(i)
            cur.wait_stream(d2h_stream)
            e1.record()
            barrier()
            ms_kp = max_over_ranks(e0.elapsed_time(e1))

    peak_tf, peak_hbm, peak_kind = measured_peaks()
    sweep = None
    if world == 1 and not args.no_lbs_sweep:
        sweep = lbs_sweep(dev, [1, 8, 64, 256, 1024, 4096], peak_hbm)
    train = None
    if args.train_steps > 0:
        train = train_bench(dev, rank, world, B, C, args.conv_mode, args.train_steps, x_dev, barrier, max_over_ranks, no_graph=args.no_train_graph)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e = world * B / (ms_e2e / args.steps * 1e-3)
    flops = algorithmic_mflop(C) * 1e6 * B
    achieved_tf = flops / (enc_ms * 1e-3) / 1e12
    cpu = None
    if world == 1:
        sample = min(B, args.cpu_sample)
        bps, sec, cores = cpu_path(sample, C, args.cpu_reps, 1)
        cpu = {'value': bps, 'unit': 'bodies/s', 'cores': cores, 'kind': 'port',
               'sample': 'median of %d passes over one B=%d batch of the same workload (oracle = CPU PyTorch '
                         'restatement of the reference path, %d threads)' % (args.cpu_reps, sample, cores)}
    traffic = conv_traffic_r02() if (B == 64 and C == 17) else None
    cfg = workload_config(B, C, world, args.conv_mode)
    line = {'metric': METRIC, 'value': value, 'unit': 'bodies/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'fp16x3 split (fp32-equivalent) convs, fp32 elsewhere' if args.conv_mode == 'f16x3_tc' else 'f32',
            'data': 'synthetic', 'config': cfg, 'conv_mode': args.conv_mode,
            'value_mode': value_mode, 'ms_per_step_eager': ms_total_eager / args.steps,
            'e2e': {'value': e2e, 'unit': 'bodies/s', 'h2d_bytes_per_step': int(x_host.numel() * 4),
                    'd2h_bytes_per_step': int(sum(h.numel() for h in out_host) * 4),
                    'h2d_copy_alone_gbs_per_rank': {'min': h2d_min, 'max': h2d_max,
                                                    'ceiling_bodies_per_s': world * B * h2d_min * 1e9 / (x_host.numel() * 4)},
                    'numa': numa},
            'e2e_from_keypoints': None if ms_kp is None else {
                'value': world * B / (ms_kp / args.steps * 1e-3), 'unit': 'bodies/s',
                'h2d_bytes_per_step': int(B * 256 * 256 * 4 + B * (C - 1) * 2 * 4), 'd2h_bytes_per_step': int(sum(h.numel() for h in out_host) * 4),
                'note': 'SingleInputRegressor.forward_from_labels: host segmentation labels + 2-D joints; the proxy representation is generated '
                        'inside the stem input pack (SURVEY 8f N1), double-buffered like e2e'},
            'gpu_launches': int(launches),      # library kernels of the timed steps (counted in the eager loop; a replay launches the same ones)
            'roofline': {'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': achieved_tf / peak_tf,
                         'traffic': traffic['dram_bytes_per_launch'] if traffic else None, 'peak_kind': peak_kind,
                         'kernel': 'encoder = pack + conv1_s2d_kernel (stem, max pool fused) + 19 conv_tc_kernel launches + avgpool',
                         'encoder_ms': enc_ms, 'algorithmic_gflop_per_step': flops / 1e9,
                         'traffic_unit': 'DRAM bytes per convolution-kernel launch (dram__bytes_read+write, mean over the 20 convolution launches '
                                         'of a step; profiles/r02_conv_traffic.json)',
                         'note': 'algorithmic FLOPs (1 pass) over the whole encoder time; the f16x3 mode issues 3 MMA passes, so 0.333 is the cap'},
            'cpu_baseline': cpu, 'train': train, 'lbs_sweep': sweep, 'clocks': clocks, 'torch': torch.__version__}
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
    ap.add_argument('--train-steps', type=int, default=10, help='timed steps of the training-step arm (0 = skip)')
    ap.add_argument('--no-lbs-sweep', action='store_true')
    ap.add_argument('--train-only', action='store_true', help='run only the training arm and print {"train": ...} (development)')
    ap.add_argument('--no-graph', action='store_true', help='time the eager loop only (ncu launch lists: replays hide the kernels)')
    ap.add_argument('--no-train-graph', action='store_true', help='time the eager training loop only (ncu launch lists: replays hide the kernels)')
    ap.add_argument('--no-numa-bind', action='store_true', help='leave the process unbound (A/B of the NUMA binding)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
