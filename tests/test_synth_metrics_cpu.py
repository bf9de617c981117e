"""CPU: SURVEY 8f rows N2 / N4 -- the oracle restatements against fixtures produced by the UNMODIFIED reference
(tests/golden/synth_b4.npz, metrics_b4.npz from oracle/gen_golden.py), the 3x3 Procrustes solve the device kernel compiles
(csrc/procrustes.h, built here for the host with g++) against numpy's SVD, and the tracker's host-side bookkeeping."""
import ctypes
import os
import pickle
import subprocess

import numpy as np
import pytest
import torch

import straps_oracle as O
from conftest import golden, checksum, rel_err, PKG, REPO, SCRATCH
from golden_inputs import (synth_inputs, metrics_inputs, SYNTH_STD, SYNTH_RANGE, SYNTH_XY_STD, SYNTH_Z_RANGE, ALL_METRICS,
                           ALL_TASKS)


# ---------------------------------------------------------------------------------------------- N2
def test_synth_oracle_matches_reference_fixture():
    g = golden('synth_b4.npz')
    I = synth_inputs()
    assert g['in_checksum'].tolist() == [checksum(v) for v in I.values()]
    T = {k: torch.from_numpy(v) for k, v in I.items()}
    R = O.batch_rodrigues(T['pose_aa'].reshape(-1, 3)).view(-1, 24, 3, 3)
    assert np.array_equal(R[:, 1:].numpy(), g['pose_rotmats']) and np.array_equal(R[:, :1].numpy(), g['glob_rotmats'])
    # the affine part is bit-exact given the draws
    n = O.sample_shape_from_noise(torch.from_numpy(g['noise_normal']), T['mean_shape'], 'normal', std_vector=torch.tensor([SYNTH_STD] * 10))
    u = O.sample_shape_from_noise(torch.from_numpy(g['noise_uniform']), T['mean_shape'], 'uniform', delta_betas_range=SYNTH_RANGE)
    assert np.array_equal(n.numpy(), g['shape_normal']) and np.array_equal(u.numpy(), g['shape_uniform'])
    c = O.cam_t_from_noise(T['mean_cam_t'], torch.from_numpy(g['noise_xy']), torch.from_numpy(g['noise_z']), SYNTH_XY_STD, SYNTH_Z_RANGE)
    assert np.array_equal(c.numpy(), g['aug_cam_t'])
    K = torch.tensor([[5000., 0., 128.], [0., 5000., 128.], [0., 0., 1.]])[None].expand(4, -1, -1)
    proj = O.perspective_project(T['points'], T['cam_R'], T['cam_t'], K)
    assert np.array_equal(proj.numpy(), g['proj']) and np.array_equal(g['proj'], g['proj_default_K'])


def test_rodrigues_zero_rotation_is_finite_identity():
    """The 1e-8 inside the norm: an exactly-zero axis-angle vector gives the identity, not NaN."""
    R = O.batch_rodrigues(torch.zeros(2, 3))
    assert torch.isfinite(R).all() and np.allclose(R.numpy(), np.eye(3)[None], atol=1e-7)


# ---------------------------------------------------------------------------------------------- N4
def test_metric_oracle_matches_reference_tracker_fixture():
    g = golden('metrics_b4.npz')
    pred, target, pr, tr_, losses = metrics_inputs()
    assert g['in_checksum'].tolist() == [checksum(pred['verts']), checksum(target['verts']), checksum(pr), checksum(tr_),
                                         checksum(pred['joints3D']), checksum(target['joints3D'])]
    ref = dict(zip(g['sum_keys'].tolist(), g['sum_values'].tolist()))
    mine = O.metric_sums(pred, target, pred_reposed=pr, target_reposed=tr_)
    assert sorted(mine) == sorted(ALL_METRICS)
    for k, v in mine.items():
        # float32 numpy on both sides; only the SVD is batched instead of looped
        assert abs(v - ref['train_' + k]) <= 2e-5 * abs(ref['train_' + k]), (k, v, ref['train_' + k])
        assert ref['val_' + k] == ref['train_' + k]
    assert rel_err(O.procrustes_batch(pred['joints3D'], target['joints3D']), g['joints3D_pa']) < 1e-5
    assert rel_err(O.scale_translation_batch(pred['joints3D'], target['joints3D']), g['joints3D_sc']) < 1e-6
    assert rel_err(O.procrustes_batch(pred['verts'][:1], target['verts'][:1])[0, :32], g['verts0_pa_first32']) < 1e-5
    B = pred['verts'].shape[0]
    assert ref['train_num_samples'] == B and abs(ref['train_losses'] - float(losses['total']) * B) < 1e-6


def test_procrustes_is_invariant_to_similarity_of_the_prediction():
    """Property (any size): aligning s.R.p + t gives the same result as aligning p."""
    rng = np.random.RandomState(3)
    p, t = rng.normal(0, 1, (5, 40, 3)), rng.normal(0, 1, (5, 40, 3))
    q, _ = np.linalg.qr(rng.normal(0, 1, (5, 3, 3)))
    q *= np.sign(np.linalg.det(q))[:, None, None]
    p2 = 1.7 * (p @ np.swapaxes(q, 1, 2)) + rng.normal(0, 1, (5, 1, 3))
    assert np.allclose(O.procrustes_batch(p, t), O.procrustes_batch(p2, t), atol=1e-10)
    # and a target that IS a similarity transform of the prediction is matched exactly
    assert np.allclose(O.procrustes_batch(p, p2), p2, atol=1e-10)


@pytest.fixture(scope='module')
def host_procrustes():
    """csrc/procrustes.h compiled for the host: the SAME source the device kernel uses."""
    os.makedirs(SCRATCH, exist_ok=True)
    so = os.path.join(SCRATCH, 'host_procrustes.so')
    subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-I', os.path.join(PKG, 'csrc'), os.path.join(REPO, 'tests', 'host_procrustes.cpp'),
                    '-o', so], check=True)
    return ctypes.CDLL(so)


def test_device_procrustes_source_matches_numpy_svd(host_procrustes):
    rng = np.random.RandomState(0)
    worst = 0.0
    for trial in range(600):
        n = rng.randint(4, 60)
        X1 = rng.normal(0, 1, (3, n))
        if trial % 3 == 0:                                   # mirrored target: det(U V^T) = -1, the Z fix engages
            X2 = X1 * np.array([[1.], [1.], [-1.]]) + 0.01 * rng.normal(0, 1, (3, n))
        else:
            X2 = rng.normal(0, 1, (3, n))
        X1 -= X1.mean(1, keepdims=True)
        X2 -= X2.mean(1, keepdims=True)
        K = X1 @ X2.T
        U, s, Vh = np.linalg.svd(K)
        Z = np.eye(3)
        Z[2, 2] = np.sign(np.linalg.det(U @ Vh))
        R = Vh.T @ Z @ U.T                                   # utils/eval_utils.py:30-38
        K9, R9 = np.ascontiguousarray(K.ravel()), np.zeros(9)
        host_procrustes.host_procrustes_rotation(K9.ctypes.data_as(ctypes.c_void_p), R9.ctypes.data_as(ctypes.c_void_p))
        worst = max(worst, float(np.abs(R9.reshape(3, 3) - R).max()))
        assert abs(np.linalg.det(R9.reshape(3, 3)) - 1.0) < 1e-10
    assert worst < 1e-10, worst


def test_device_jacobi_source_diagonalises(host_procrustes):
    rng = np.random.RandomState(1)
    for _ in range(200):
        M = rng.normal(0, 1, (3, 3))
        A = M @ M.T
        A9, V9, w3 = np.ascontiguousarray(A.ravel()), np.zeros(9), np.zeros(3)
        host_procrustes.host_jacobi_sym3(A9.ctypes.data_as(ctypes.c_void_p), V9.ctypes.data_as(ctypes.c_void_p),
                                         w3.ctypes.data_as(ctypes.c_void_p))
        V = V9.reshape(3, 3)
        assert np.allclose(V @ np.diag(w3) @ V.T, A, atol=1e-12 * max(1.0, np.abs(A).max()))
        assert np.allclose(sorted(w3), np.linalg.eigvalsh(A), atol=1e-12 * max(1.0, np.abs(A).max()))


# ---------------------------------------------------------------------------------------------- tracker bookkeeping (host)
def test_tracker_epoch_bookkeeping_matches_reference_fixture(tmp_path):
    """update_per_epoch / history / pickle log with the reference's per-batch sums injected: same history values."""
    from metrics.train_loss_and_metrics_tracker import TrainingLossesAndMetricsTracker
    g = golden('metrics_b4.npz')
    log = str(tmp_path / 'log.pkl')
    tr = TrainingLossesAndMetricsTracker(ALL_TASKS, ALL_METRICS, 256, log)
    assert sorted(tr.history) == g['history_keys'].tolist()
    tr.initialise_loss_metric_sums()
    assert sorted(tr.loss_metric_sums) == g['sum_keys'].tolist()
    tr.loss_metric_sums.update(dict(zip(g['sum_keys'].tolist(), g['sum_values'].tolist())))
    tr.update_per_epoch()
    for k, v in zip(g['history_keys'].tolist(), g['history_values'].tolist()):
        # the reference's sums are numpy float32 scalars (NEP 50), so its division rounds to float32; ours is float64
        assert tr.history[k][-1] == pytest.approx(v, rel=1e-6), k
    with open(log, 'rb') as f:
        assert pickle.load(f) == tr.history
    # resume: cut to the current epoch, zero-fill what an old log lacks
    with open(log, 'wb') as f:
        pickle.dump({k: v for k, v in tr.history.items() if 'pve-ts' not in k}, f)
    tr2 = TrainingLossesAndMetricsTracker(ALL_TASKS, ALL_METRICS, 256, log, load_logs=True, current_epoch=1)
    assert tr2.history['train_pve-ts'] == [0.0] and tr2.history['val_pves'] == tr.history['val_pves']
    best = {m: tr.history['val_' + m][-1] for m in ('pves', 'mpjpes')}
    assert tr.determine_save_model_weights_this_epoch(['pves', 'mpjpes'], best)
    assert not tr.determine_save_model_weights_this_epoch(['pves'], {'pves': best['pves'] * 0.5})


def test_tracker_untracked_series_follow_the_reference(tmp_path):
    """Untracked task losses get a 0 per epoch; untracked metrics get nothing (reference update_per_epoch)."""
    from metrics.train_loss_and_metrics_tracker import TrainingLossesAndMetricsTracker
    tr = TrainingLossesAndMetricsTracker(['verts'], ['pves'], 256, str(tmp_path / 'log.pkl'))
    tr.initialise_loss_metric_sums()
    tr.loss_metric_sums.update({'train_num_samples': 2, 'val_num_samples': 4, 'train_losses': 3.0, 'val_losses': 2.0,
                                'train_verts_losses': 1.0, 'val_verts_losses': 8.0, 'train_pves': 6890.0, 'val_pves': 4 * 6890.0})
    tr.update_per_epoch()
    assert tr.history['train_losses'] == [1.5] and tr.history['val_verts_losses'] == [2.0]
    assert tr.history['train_joints2D_losses'] == [0.] and tr.history['train_mpjpes'] == []
    assert tr.history['train_pves'] == [0.5] and tr.history['val_pves'] == [1.0]
    with pytest.raises(Exception):          # CPU tensors: no fallback
        tr.update_per_batch('train', torch.tensor(1.0), {'verts': torch.tensor(1.0)}, {'verts': torch.zeros(1, 4, 3)},
                            {'verts': torch.zeros(1, 4, 3)}, 1)
