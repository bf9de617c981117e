"""BASELINE config 4: the training step sharded data-parallel over N GPUs with ONE NCCL all-reduce on the flat gradient bucket.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dp_check.py [--batch 64] [--steps 5] [--oracle-batch 0]

Every rank: identical initial weights (broadcast), its own seeded shard, forward + loss + backward through the drop-in API,
then DataParallelAdam.step() = all-reduce(sum) of the 11.9 M-element bucket + fused Adam with 1/N folded in.
Checks: (1) the reduced bucket is bit-identical on all ranks and equals the sum of the per-rank gradients; (2) parameters are
bit-identical on all ranks after the update; (3) optionally (--oracle-batch B>0, small B) the averaged gradient against the mean
of the CPU oracle's per-shard gradients.  Prints one JSON line with the training throughput (bodies/s over all ranks).
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'oracle'))
sys.path.insert(0, os.path.join(REPO, 'straps-3dhumanshapepose_b200'))
os.environ.setdefault('STRAPS_ASSETS_ROOT', os.path.join(REPO, 'tests', '_scratch', 'assets'))
import numpy as np
import torch
import torch.distributed as dist

TASKS = ['verts', 'joints2D', 'joints3D', 'pose_params', 'shape_params']
W = {'verts': 1.0, 'joints2D': 0.1, 'pose_params': 0.1, 'shape_params': 0.1, 'joints3D': 1.0}


def make_targets(B, seed, smpl_oracle, O):
    rng = np.random.RandomState(seed)
    betas = torch.from_numpy(rng.normal(0, 1, (B, 10)).astype(np.float32))
    with torch.no_grad():
        R = O.rot6d_to_rotmat(torch.from_numpy(rng.normal(0, 1, (B, 144)).astype(np.float32))).view(B, 24, 3, 3)
        v, j = smpl_oracle.forward_rotmats(R, betas)
    j2d = torch.from_numpy(rng.uniform(-20, 276, (B, 17, 2)).astype(np.float32))
    return {'verts': v, 'joints2D': j2d, 'joints3D': j[:, O.ALL_JOINTS_TO_H36M_MAP][:, O.H36M_TO_J14], 'shape_params': betas,
            'pose_params_rot_matrices': R}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64, help='per-GPU batch')
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--channels', type=int, default=17)
    ap.add_argument('--oracle-batch', type=int, default=0)
    ap.add_argument('--overlap', action='store_true', help='also check and time the early all-reduce of layer4 + IEF inside the backward pass')
    ap.add_argument('--conv-mode', default='f16x3_tc', choices=['fp32_simt', 'f16x3_tc'])
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import straps_oracle as O
    from straps_b200 import synthetic_assets, synthetic_inputs
    from straps_b200.parallel import DataParallelAdam
    if rank == 0:
        synthetic_assets.write_synthetic_assets(os.environ['STRAPS_ASSETS_ROOT'], seed=0)
    if world > 1:
        dist.barrier()
    import config
    from models.regressor import SingleInputRegressor
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat
    from utils.cam_utils import orthographic_project_torch
    from utils.joints2d_utils import check_joints2d_visibility_torch
    from losses.multi_task_loss import HomoscedasticUncertaintyWeightedMultiTaskLoss as Loss
    add = os.path.join(os.environ['STRAPS_ASSETS_ROOT'], 'additional')
    smpl_oracle = O.SmplOracle(add, batch_size=args.batch)
    B, C = args.batch, args.channels
    torch.manual_seed(1234 + rank)          # deliberately different: the broadcast in DataParallelAdam must equalise the replicas
    reg = SingleInputRegressor(C, 18, 3, conv_mode=args.conv_mode).to(dev).train()
    crit = Loss(TASKS, init_loss_weights=W).to(dev)
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(dev)
    params = [p for n, p in reg.named_parameters()] + list(crit.parameters())
    opt = DataParallelAdam(params, lr=1e-4)
    assert opt.bucket.numel == (11906658 if C == 17 else 11909794), opt.bucket.numel

    def step_inputs(step):
        x = torch.from_numpy(synthetic_inputs.make_proxy_batch(B, C, seed=1000 * step + rank)).to(dev)
        tg = {k: v.to(dev) for k, v in make_targets(B, 7000 * step + rank, smpl_oracle, O).items()}
        tg['vis'] = check_joints2d_visibility_torch(tg['joints2D'], config.REGRESSOR_IMG_WH)
        return x, tg

    def forward_backward(x, tg):
        opt.zero_grad()
        cam, pose, shape = reg(x)
        R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
        outs = {'verts': out.vertices, 'joints2D': orthographic_project_torch(out.joints[:, config.ALL_JOINTS_TO_COCO_MAP, :], cam),
                'joints3D': out.joints[:, config.ALL_JOINTS_TO_H36M_MAP, :][:, config.H36M_TO_J14, :], 'shape_params': shape,
                'pose_params_rot_matrices': R}
        loss, _ = crit(tg, outs)
        loss.backward()
        return loss

    # ---- correctness of the exchange on step 0
    x, tg = step_inputs(0)
    loss = forward_backward(x, tg)
    opt.bucket.gather()
    local_grads = opt.bucket.grads.clone()
    opt.all_reduce()
    reduced = opt.bucket.grads.clone()
    ok = {}
    if world > 1:
        all_red = [torch.empty_like(reduced) for _ in range(world)]
        all_loc = [torch.empty_like(reduced) for _ in range(world)]
        dist.all_gather(all_red, reduced)
        dist.all_gather(all_loc, local_grads)
        ok['reduced_bit_identical'] = all(torch.equal(all_red[0], t) for t in all_red)
        s = torch.stack(all_loc).double().sum(0)
        ok['reduced_vs_sum_rel_err'] = float((reduced.double() - s).abs().max() / s.abs().max())
    if args.overlap and world > 1:
        # the same exchange with the layer4 + IEF part started inside the backward pass on a side stream (DataParallelAdam.enable_overlap):
        # same inputs, same weights -> the reduced bucket must agree with the plain one (to the run-to-run noise of the atomics)
        ok['overlap_enabled'] = bool(opt.enable_overlap(reg))
        forward_backward(x, tg)
        opt.all_reduce()
        ok['overlap_reduced_rel_err'] = float((opt.bucket.grads.double() - reduced.double()).abs().max() / reduced.double().abs().max())
        all_red2 = [torch.empty_like(reduced) for _ in range(world)]
        dist.all_gather(all_red2, opt.bucket.grads.clone())
        ok['overlap_reduced_bit_identical_across_ranks'] = all(torch.equal(all_red2[0], t) for t in all_red2)
        opt.apply_update()
    else:
        opt.bucket.grads.copy_(local_grads)          # step() performs the all-reduce itself
        opt.step()
    if world > 1:
        allp = [torch.empty_like(opt.bucket.params) for _ in range(world)]
        dist.all_gather(allp, opt.bucket.params)
        ok['params_bit_identical_after_step'] = all(torch.equal(allp[0], t) for t in allp)
    # ---- timing: forward + loss + backward + all-reduce + Adam
    for s in range(1, 3):
        forward_backward(*step_inputs(s)); opt.step()
    data = [step_inputs(10 + s) for s in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for x, tg in data:
        forward_backward(x, tg)
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({'config': 'BASELINE config %d: training step (encoder+IEF+SMPL+multi-task loss, conv mode %s), '
                                    'B=%d/GPU, %d GPU(s), one all-reduce of %d fp32 + fused Adam' % (4 if world > 1 else 3, args.conv_mode, B, world, opt.bucket.numel),
                          'ms_per_step': ms, 'bodies_per_s': world * B / (ms * 1e-3), 'loss0': float(loss), 'checks': ok}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
