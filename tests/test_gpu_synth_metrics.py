"""GPU parity, SURVEY 8f rows N2 / N4: the target-side synthesis kernels (csrc/synth.cu) and the device-side metrics
(csrc/metrics.cu) through the drop-in modules (augmentation/, utils/cam_utils.py, utils/eval_utils.py, metrics/) -> ctypes ->
C ABI, against fixtures produced by the unmodified reference, the CPU oracle at other sizes, and size-independent
properties.  Tolerances: bit-exact for the affine augmentation given the same draws; 1e-5 relative for the rotations /
projection; 1e-4 relative for the metric sums (north_star bar; measured errors are ~1e-6)."""
import numpy as np
import pytest
import torch

import straps_oracle as O
from conftest import golden, rel_err, RTOL
from golden_inputs import (synth_inputs, metrics_inputs, SYNTH_STD, SYNTH_RANGE, SYNTH_XY_STD, SYNTH_Z_RANGE, ALL_METRICS,
                           ALL_TASKS)

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# ---------------------------------------------------------------------------------------------- N2
def test_rodrigues_and_perspective_against_reference_fixture():
    from straps_b200 import ops
    from utils.cam_utils import perspective_project_torch
    g = golden('synth_b4.npz')
    I = synth_inputs()
    R = ops.batch_rodrigues(cu(I['pose_aa']).reshape(-1, 3)).view(-1, 24, 3, 3).cpu().numpy()
    assert rel_err(R[:, 1:], g['pose_rotmats']) < 1e-5 and rel_err(R[:, :1], g['glob_rotmats']) < 1e-5
    assert np.abs(R[0, 1] - np.eye(3)).max() < 1e-6          # the exactly-zero rotation of the fixture
    K = torch.tensor([[5000., 0., 128.], [0., 5000., 128.], [0., 0., 1.]], device=DEV)[None].expand(4, -1, -1)
    proj = perspective_project_torch(cu(I['points']), cu(I['cam_R']), cu(I['cam_t']), cam_K=K)
    assert proj.shape == (4, 17, 2) and rel_err(proj.cpu().numpy(), g['proj']) < 1e-5
    proj2 = perspective_project_torch(cu(I['points']), cu(I['cam_R']), cu(I['cam_t']), focal_length=5000., img_wh=256)
    assert torch.equal(proj, proj2)


def test_augmentation_affine_part_is_bit_exact_given_the_draws():
    from straps_b200 import ops
    g = golden('synth_b4.npz')
    I = synth_inputs()
    mean = cu(I['mean_shape'])
    n = ops.scale_shift(cu(g['noise_normal']), torch.tensor([SYNTH_STD] * 10, device=DEV), None, mean)
    u = ops.scale_shift(cu(g['noise_uniform']), SYNTH_RANGE[1] - SYNTH_RANGE[0], SYNTH_RANGE[0], mean)
    assert np.array_equal(n.cpu().numpy(), g['shape_normal']) and np.array_equal(u.cpu().numpy(), g['shape_uniform'])
    noise = torch.cat([cu(g['noise_xy']), cu(g['noise_z'])[:, None]], dim=1)
    l, h = SYNTH_Z_RANGE
    c = ops.scale_shift(noise, torch.tensor([SYNTH_XY_STD, SYNTH_XY_STD, h - l], device=DEV), torch.tensor([0., 0., l], device=DEV),
                        cu(I['mean_cam_t'])[:1].expand(4, -1))       # expanded (stride 0) like run_train.py:116
    assert np.array_equal(c.cpu().numpy(), g['aug_cam_t'])


@pytest.mark.parametrize('dist', ['normal', 'uniform'])
def test_augment_drop_ins_consume_the_generator_like_the_reference(dist):
    """Same CUDA seed => the drop-ins equal the reference's torch-op sequence (restated by the oracle functions, evaluated with
    torch's own CUDA ops) bit for bit: same draw shapes, order and rounding."""
    from augmentation.smpl_augmentation import augment_smpl
    from augmentation.cam_augmentation import augment_cam_t
    B = 64
    rng = np.random.RandomState(5)
    pose = cu(rng.normal(0, 0.5, (B, 72)).astype(np.float32))
    orig = cu(rng.normal(0, 1, (B, 10)).astype(np.float32))
    mean_shape = cu(rng.normal(0, 0.3, 10).astype(np.float32))
    mean_cam_t = torch.tensor([0., 0.2, 42.], device=DEV)[None].expand(B, -1)
    std_vec = torch.tensor([SYNTH_STD] * 10, device=DEV)
    params = {'augment_shape': True, 'delta_betas_distribution': dist, 'delta_betas_range': SYNTH_RANGE, 'delta_betas_std_vector': std_vec}
    torch.manual_seed(123)
    shape, pose_rm, glob_rm = augment_smpl(orig, pose[:, 3:], pose[:, :3], mean_shape, params)
    cam_t = augment_cam_t(mean_cam_t, xy_std=SYNTH_XY_STD, delta_z_range=SYNTH_Z_RANGE)
    torch.manual_seed(123)
    noise = torch.randn(B, 10, device=DEV) if dist == 'normal' else torch.rand(B, 10, device=DEV)
    exp_shape = O.sample_shape_from_noise(noise, mean_shape, dist, delta_betas_range=SYNTH_RANGE, std_vector=std_vec)
    exp_cam = O.cam_t_from_noise(mean_cam_t, torch.randn(B, 2, device=DEV), torch.rand(B, device=DEV), SYNTH_XY_STD, SYNTH_Z_RANGE)
    assert torch.equal(shape, exp_shape) and torch.equal(cam_t, exp_cam)
    assert pose_rm.shape == (B, 23, 3, 3) and glob_rm.shape == (B, 1, 3, 3)
    R = torch.cat([glob_rm, pose_rm], dim=1)
    assert rel_err(R.cpu().numpy(), O.batch_rodrigues(pose.cpu().reshape(-1, 3)).view(B, 24, 3, 3).numpy()) < 1e-5
    eye = torch.eye(3, device=DEV)
    assert float((R @ R.transpose(-1, -2) - eye).abs().max()) < 1e-5       # property: rotations are orthonormal
    # augment_shape off: the shape passes through untouched
    params['augment_shape'] = False
    assert augment_smpl(orig, pose[:, 3:], pose[:, :3], mean_shape, params)[0] is orig


def test_target_side_of_the_training_loop(assets_root, smpl_oracle):
    """train/train_synthetic_otf_rendering.py:121-145 end to end on the device: augmentation -> target SMPL (rotation
    matrices) -> COCO joints -> perspective projection -> T-pose SMPL, against the CPU oracle on the same draws."""
    import config
    from models.smpl_official import SMPL
    from augmentation.smpl_augmentation import augment_smpl
    from utils.cam_utils import perspective_project_torch
    B = 8
    rng = np.random.RandomState(9)
    pose = cu(rng.normal(0, 0.4, (B, 72)).astype(np.float32))
    shape0 = cu(rng.normal(0, 1, (B, 10)).astype(np.float32))
    mean_shape = cu(np.zeros(10, np.float32))
    params = {'augment_shape': True, 'delta_betas_distribution': 'normal', 'delta_betas_range': SYNTH_RANGE,
              'delta_betas_std_vector': torch.tensor([SYNTH_STD] * 10, device=DEV)}
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(DEV)
    cam_t = torch.tensor([0., 0.2, 42.], device=DEV)[None].expand(B, -1)
    cam_R = torch.eye(3, device=DEV)[None].expand(B, -1, -1)
    K = torch.tensor([[5000., 0., 128.], [0., 5000., 128.], [0., 0., 1.]], device=DEV)[None].expand(B, -1, -1)
    with torch.no_grad():
        torch.manual_seed(3)
        shape, pose_rm, glob_rm = augment_smpl(shape0, pose[:, 3:], pose[:, :3], mean_shape, params)
        out = smpl(body_pose=pose_rm, global_orient=glob_rm, betas=shape, pose2rot=False)
        j2d = perspective_project_torch(out.joints[:, config.ALL_JOINTS_TO_COCO_MAP, :], cam_R, cam_t, cam_K=K)
        reposed = smpl(betas=shape).vertices
        # oracle on the same augmented shape
        R = O.batch_rodrigues(pose.cpu().reshape(-1, 3)).view(B, 24, 3, 3)
        v, j = smpl_oracle.forward_rotmats(R, shape.cpu())
        ej2d = O.perspective_project(j[:, O.ALL_JOINTS_TO_COCO_MAP, :], cam_R.cpu(), cam_t.cpu(), K.cpu())
        vt, _ = smpl_oracle.forward(betas=shape.cpu(), body_pose=torch.zeros(B, 69), global_orient=torch.zeros(B, 3), pose2rot=True)
    assert rel_err(out.vertices.cpu().numpy(), v.numpy()) < RTOL and rel_err(out.joints.cpu().numpy(), j.numpy()) < RTOL
    assert rel_err(j2d.cpu().numpy(), ej2d.numpy()) < RTOL
    assert rel_err(reposed.cpu().numpy(), vt.numpy()) < RTOL


# ---------------------------------------------------------------------------------------------- N4
def _expected_sums():
    g = golden('metrics_b4.npz')
    return dict(zip(g['sum_keys'].tolist(), g['sum_values'].tolist())), g


def test_tracker_sums_against_reference_fixture(tmp_path):
    from metrics.train_loss_and_metrics_tracker import TrainingLossesAndMetricsTracker
    ref, g = _expected_sums()
    pred, target, pr, tr_, losses = metrics_inputs()
    B = pred['verts'].shape[0]
    t = lambda d: {k: cu(v) for k, v in d.items()}
    tracker = TrainingLossesAndMetricsTracker(ALL_TASKS, ALL_METRICS, 256, str(tmp_path / 'log.pkl'))
    tracker.initialise_loss_metric_sums()
    pd, td = t(pred), t(target)
    for split in ('train', 'val'):
        tracker.update_per_batch(split, torch.tensor(float(losses['total']), device=DEV),
                                 {k: torch.tensor(float(losses[k]), device=DEV) for k in ALL_TASKS}, pd, td, B,
                                 pred_reposed_vertices=cu(pr), target_reposed_vertices=cu(tr_))
    assert torch.is_tensor(pd['verts']) and pd['verts'].is_cuda          # inputs are left alone (and on the device)
    sums = tracker.sync()
    assert sorted(sums) == sorted(ref)
    for k, v in ref.items():
        assert abs(sums[k] - v) <= RTOL * max(abs(v), 1e-12), (k, sums[k], v)
    tracker.update_per_epoch()
    for k, v in zip(g['history_keys'].tolist(), g['history_values'].tolist()):
        assert tracker.history[k][-1] == pytest.approx(v, rel=RTOL), k
    # a second epoch starts from zero
    tracker.initialise_loss_metric_sums()
    assert all(v == 0 for v in tracker.sync().values())


def test_eval_utils_drop_ins_against_reference_fixture():
    from utils.eval_utils import procrustes_analysis_batch, scale_and_translation_transform_batch, compute_similarity_transform
    _, g = _expected_sums()
    pred, target, _, _, _ = metrics_inputs()
    pa = procrustes_analysis_batch(cu(pred['joints3D']), cu(target['joints3D']))
    sc = scale_and_translation_transform_batch(cu(pred['joints3D']), cu(target['joints3D']))
    assert pa.is_cuda and rel_err(pa.cpu().numpy(), g['joints3D_pa']) < 1e-5
    assert rel_err(sc.cpu().numpy(), g['joints3D_sc']) < 1e-5
    # numpy in -> numpy out like the reference (computed on the device all the same); [N,3] and [3,N] forms
    one = compute_similarity_transform(pred['verts'][0], target['verts'][0])
    assert isinstance(one, np.ndarray) and one.dtype == np.float32 and rel_err(one[:32], g['verts0_pa_first32']) < 1e-5
    one_t = compute_similarity_transform(pred['verts'][0].T, target['verts'][0].T)
    assert one_t.shape == (3, 6890) and np.array_equal(one_t.T, one)


@pytest.mark.parametrize('B,N', [(1, 3), (3, 14), (64, 6890), (257, 90)])
def test_point_metrics_against_oracle_and_properties(B, N):
    """Other sizes (ragged w.r.t. the 256-thread CTA, N < warp, BASELINE's B=64 x 6890) against the numpy oracle, plus
    size-independent properties of the alignment."""
    from straps_b200 import ops
    rng = np.random.RandomState(B * 1000 + N)
    p = rng.normal(0, 0.3, (B, N, 3)).astype(np.float32)
    q, _ = np.linalg.qr(rng.normal(0, 1, (B, 3, 3)))
    q *= np.sign(np.linalg.det(q))[:, None, None]
    t = (1.3 * (p @ np.swapaxes(q, 1, 2)) + rng.normal(0, 0.2, (B, 1, 3)) + rng.normal(0, 0.02, (B, N, 3))).astype(np.float32)
    P, T = cu(p), cu(t)
    sums = torch.zeros(3, dtype=torch.float64, device=DEV)
    sc, pa = ops.points_metrics(P, T, 7, sums=sums, want_sc=True, want_pa=True)
    e_sc, e_pa = O.scale_translation_batch(p, t), O.procrustes_batch(p.astype(np.float64), t.astype(np.float64))
    assert rel_err(sc.cpu().numpy(), e_sc) < 1e-5 and rel_err(pa.cpu().numpy(), e_pa) < 1e-5
    exp = [np.linalg.norm(p - t, axis=-1).sum(dtype=np.float64), np.linalg.norm(e_sc - t, axis=-1).sum(dtype=np.float64),
           np.linalg.norm(e_pa - t, axis=-1).sum(dtype=np.float64)]
    got = sums.cpu().numpy()
    for a, b in zip(got, exp):
        assert abs(a - b) <= RTOL * b, (got, exp)
    # selector bits: only the requested slots are touched, and sums ACCUMULATE
    ops.points_metrics(P, T, ops.METRIC_PA, sums=sums)
    again = sums.cpu().numpy()
    assert again[0] == got[0] and again[1] == got[1] and abs(again[2] - 2 * got[2]) <= 1e-9 * got[2]
    # property: the Procrustes result does not change when the prediction is moved by a similarity transform
    p2 = cu((0.6 * (p @ np.swapaxes(q, 1, 2)[::-1]) + 0.5).astype(np.float32))
    _, pa2 = ops.points_metrics(p2, T, ops.METRIC_PA, want_pa=True)
    assert rel_err(pa2.cpu().numpy(), pa.cpu().numpy()) < 1e-4
    # property: the scale-and-translation correction is idempotent, and exact when pred == target
    sc2, _ = ops.points_metrics(sc, T, ops.METRIC_SC, want_sc=True)
    assert rel_err(sc2.cpu().numpy(), sc.cpu().numpy()) < 1e-5
    same, pa_same = ops.points_metrics(T, T, 6, want_sc=True, want_pa=True)
    assert rel_err(same.cpu().numpy(), t) < 1e-5 and rel_err(pa_same.cpu().numpy(), t) < 1e-5


def test_metric_kernels_reject_bad_arguments():
    from straps_b200 import ops
    from straps_b200._lib import StrapsError
    a = torch.zeros(2, 5, 3, device=DEV)
    with pytest.raises(StrapsError):
        ops.points_metrics(a.cpu(), a.cpu(), 1, sums=torch.zeros(3, dtype=torch.float64, device=DEV))
    with pytest.raises(StrapsError):
        ops.points_metrics(a, a, 1, sums=torch.zeros(3, dtype=torch.float32, device=DEV))
    with pytest.raises(StrapsError):
        ops.points_metrics(a, a, 0, sums=torch.zeros(3, dtype=torch.float64, device=DEV))
    with pytest.raises(StrapsError):
        ops.batch_rodrigues(torch.zeros(4, 4, device=DEV))
    ops.points_metrics(a[:0], a[:0], 1, sums=torch.zeros(3, dtype=torch.float64, device=DEV))      # empty batch: no launch, no error
