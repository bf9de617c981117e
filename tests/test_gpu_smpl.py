"""GPU parity: fused SMPL forward / rot6d / projection kernels (through the drop-in API -> ctypes -> C ABI)
against the CPU oracle and the reference-generated fixtures.  Tolerance: 1e-4 relative fp32 (north_star),
bit-exact for index work (vertex picks, faces, parents)."""
import numpy as np
import pytest
import torch

import straps_oracle as O
from conftest import golden, smpl_inputs, rel_err, RTOL

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


_MODELS = {}


def _smpl(batch):
    """SMPL default parameters are sized by the constructor's batch_size."""
    import config
    from models.smpl_official import SMPL
    if batch not in _MODELS:
        _MODELS[batch] = SMPL(config.SMPL_MODEL_DIR, batch_size=batch).to(DEV)
    return _MODELS[batch]


@pytest.fixture(scope='module')
def smpl(assets_root):
    return _smpl(4)


def _oracle_rotmats(smpl_oracle, pose6d, betas):
    with torch.no_grad():
        R = O.rot6d_to_rotmat(torch.from_numpy(pose6d)).view(-1, 24, 3, 3)
        v, j = smpl_oracle.forward_rotmats(R, torch.from_numpy(betas))
    return R, v, j


def test_config1_b4_against_reference_fixture(smpl):
    """BASELINE config 1: SMPL forward only, B=4, random pose/shape."""
    from utils.rigid_transform_utils import rot6d_to_rotmat
    g = golden('smpl_b4.npz')
    betas, pose6d, aa = smpl_inputs(4, 11)
    with torch.no_grad():
        R = rot6d_to_rotmat(torch.from_numpy(pose6d).to(DEV)).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=torch.from_numpy(betas).to(DEV), pose2rot=False)
        out_aa = smpl(body_pose=torch.from_numpy(aa[:, 3:]).to(DEV), global_orient=torch.from_numpy(aa[:, :3]).to(DEV),
                      betas=torch.from_numpy(betas).to(DEV))
        out_t = smpl(betas=torch.from_numpy(betas).to(DEV))
    assert rel_err(R.cpu().numpy(), g['rotmats']) < 1e-5
    assert rel_err(out.vertices.cpu().numpy(), g['vertices']) < RTOL
    assert rel_err(out.joints.cpu().numpy(), g['joints']) < RTOL
    assert rel_err(out_aa.vertices.cpu().numpy(), g['vertices_aa']) < RTOL
    assert rel_err(out_aa.joints.cpu().numpy(), g['joints_aa']) < RTOL
    assert rel_err(out_t.vertices.cpu().numpy(), g['vertices_tpose']) < RTOL
    assert rel_err(out_t.joints.cpu().numpy(), g['joints_tpose']) < RTOL
    # integer work is bit exact
    assert np.array_equal(smpl.faces_tensor.cpu().numpy(), g['faces'])
    assert np.array_equal(smpl.parents.cpu().numpy(), g['parents'])
    assert out.full_pose is None and out.betas.shape == (4, 10) and out.joints.shape == (4, 90, 3)


@pytest.mark.parametrize('B', [1, 2, 3, 5, 8, 13, 33, 64, 100])
def test_ragged_batches_against_oracle(assets_root, smpl_oracle, B):
    smpl = _smpl(B)
    betas, pose6d, aa = smpl_inputs(B, 100 + B)
    R, v, j = _oracle_rotmats(smpl_oracle, pose6d, betas)
    Rg = R.to(DEV)
    with torch.no_grad():
        out = smpl(body_pose=Rg[:, 1:], global_orient=Rg[:, 0].unsqueeze(1), betas=torch.from_numpy(betas).to(DEV), pose2rot=False)
    assert rel_err(out.vertices.cpu().numpy(), v.numpy()) < RTOL
    assert rel_err(out.joints.cpu().numpy(), j.numpy()) < RTOL
    # the 21 picked joints are plain copies of vertices: bit exact against the kernel's own vertices
    idx = smpl.extra_joints_idxs
    assert torch.equal(out.joints[:, 24:45], out.vertices[:, idx])
    # consumers' index selections (config.py:27-32) are exact gathers
    import config
    coco = out.joints[:, config.ALL_JOINTS_TO_COCO_MAP]
    assert torch.equal(coco[:, 0], out.joints[:, 24]) and coco.shape == (B, 17, 3)


def test_axis_angle_and_transl(assets_root, smpl_oracle):
    B = 6
    smpl = _smpl(B)
    betas, _, aa = smpl_inputs(B, 7)
    transl = np.random.RandomState(5).normal(0, 1, (B, 3)).astype(np.float32)
    with torch.no_grad():
        v, j = smpl_oracle.forward(betas=torch.from_numpy(betas), body_pose=torch.from_numpy(aa[:, 3:]),
                                   global_orient=torch.from_numpy(aa[:, :3]), pose2rot=True)
        out = smpl(betas=torch.from_numpy(betas).to(DEV), body_pose=torch.from_numpy(aa[:, 3:]).to(DEV),
                   global_orient=torch.from_numpy(aa[:, :3]).to(DEV), transl=torch.from_numpy(transl).to(DEV))
    t = torch.from_numpy(transl)[:, None]
    assert rel_err(out.vertices.cpu().numpy(), (v + t).numpy()) < RTOL
    assert rel_err(out.joints.cpu().numpy(), (j + t).numpy()) < RTOL


def test_zero_rotation_vector_has_no_nan(smpl):
    """batch_rodrigues adds 1e-8 inside the norm, so an all-zero pose is finite (T-pose)."""
    with torch.no_grad():
        out = smpl(betas=torch.zeros(4, 10, device=DEV))
    assert torch.isfinite(out.vertices).all() and torch.isfinite(out.joints).all()
    assert rel_err(out.vertices[0].cpu().numpy(), smpl.v_template.cpu().numpy()) < 1e-6


def test_batch_size_mismatch_is_rejected_like_the_reference(smpl):
    from straps_b200._lib import StrapsError
    with pytest.raises(StrapsError):
        with torch.no_grad():
            smpl(betas=torch.zeros(9, 10, device=DEV), body_pose=torch.zeros(9, 69, device=DEV),
                 global_orient=torch.zeros(9, 3, device=DEV))


def test_properties_at_large_batch(assets_root):
    """Size-independent properties at B=256 (BASELINE config 5 sizes): identity pose -> shape blend; rigid root."""
    B = 256
    smpl = _smpl(B)
    rng = np.random.RandomState(8)
    betas = torch.from_numpy(rng.normal(0, 1, (B, 10)).astype(np.float32)).to(DEV)
    R = torch.eye(3, device=DEV).repeat(B, 24, 1, 1)
    with torch.no_grad():
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, :1], betas=betas, pose2rot=False)
        v_shaped = smpl.v_template + torch.einsum('bl,mkl->bmk', betas, smpl.shapedirs)
        J = torch.einsum('bik,ji->bjk', v_shaped, smpl.J_regressor)
    assert rel_err(out.vertices.cpu().numpy(), v_shaped.cpu().numpy()) < 1e-5
    assert rel_err(out.joints[:, :24].cpu().numpy(), J.cpu().numpy()) < 1e-5
    from utils.rigid_transform_utils import rot6d_to_rotmat
    R0 = rot6d_to_rotmat(torch.from_numpy(rng.normal(0, 1, (B, 6)).astype(np.float32)).to(DEV))
    R2 = R.clone()
    R2[:, 0] = R0
    with torch.no_grad():
        out2 = smpl(body_pose=R2[:, 1:], global_orient=R2[:, :1], betas=betas, pose2rot=False)
        expect = torch.einsum('bij,bvj->bvi', R0, v_shaped - J[:, :1]) + J[:, :1]
    assert rel_err(out2.vertices.cpu().numpy(), expect.cpu().numpy()) < 1e-5


def test_rot6d_and_projection(smpl_oracle):
    from utils.rigid_transform_utils import rot6d_to_rotmat
    from utils.cam_utils import orthographic_project_torch
    rng = np.random.RandomState(4)
    x = rng.normal(0, 1, (64, 144)).astype(np.float32)
    R = rot6d_to_rotmat(torch.from_numpy(x).to(DEV))
    Ro = O.rot6d_to_rotmat(torch.from_numpy(x))
    assert R.shape == (64 * 24, 3, 3)
    assert rel_err(R.cpu().numpy(), Ro.numpy()) < 1e-5
    assert (torch.linalg.det(R.cpu()) - 1).abs().max() < 1e-5
    pts = rng.normal(0, 1, (7, 17, 3)).astype(np.float32)
    params = rng.normal(0, 1, (7, 157)).astype(np.float32)
    cam_view = torch.from_numpy(params).to(DEV)[:, :3]            # non-contiguous view, like the IEF output
    out = orthographic_project_torch(torch.from_numpy(pts).to(DEV), cam_view)
    ref = O.orthographic_project(torch.from_numpy(pts), torch.from_numpy(params)[:, :3])
    assert rel_err(out.cpu().numpy(), ref.numpy()) < 1e-6


def test_proxy_representation_synthesis_bit_exact(assets_root):
    """SURVEY.md 8f row N1 (utils/label_conversions.py:48-55, 90-127): bit-exact against the reference fixture and the oracle."""
    from utils.label_conversions import convert_2Djoints_to_gaussian_heatmaps_torch, convert_multiclass_to_binary_labels_torch
    g = golden('heatmaps_b4.npz')
    hm = convert_2Djoints_to_gaussian_heatmaps_torch(torch.from_numpy(g['joints2d']).to(DEV), 256)
    assert hm.shape == (4, 17, 256, 256) and np.array_equal(hm.cpu().numpy(), g['heatmaps'])
    rng = np.random.RandomState(3)
    j = torch.from_numpy(rng.uniform(-20, 280, (64, 17, 2)).astype(np.float32))
    assert torch.equal(convert_2Djoints_to_gaussian_heatmaps_torch(j.to(DEV), 256).cpu(), O.joints2d_to_heatmaps(j, 256))
    seg = torch.from_numpy(rng.randint(0, 7, (5, 256, 256)).astype(np.float32))
    b = convert_multiclass_to_binary_labels_torch(seg.to(DEV))
    assert torch.equal(b.cpu(), O.binary_labels(seg)) and b.dtype == torch.float32
    # the reference assembles the regressor input like this (train/...:178-182)
    x = torch.cat([b.unsqueeze(1)[:4], hm], dim=1)
    assert x.shape == (4, 18, 256, 256)


def _lbs_mode(mode):
    """STRAPS_LBS is read by the library on every SMPL forward: 'simt' = the CUDA-core kernel, 'tc' = the tensor-core kernels at every
    batch size, unset = tensor cores from the batch size where they win."""
    import os
    os.environ.pop('STRAPS_LBS_V', None)
    if mode is None:
        os.environ.pop('STRAPS_LBS', None)
    elif mode == 'tc2':                                     # the one-CTA-per-item tensor-core kernel (lbs_tc_kernel) instead of the persistent one
        os.environ['STRAPS_LBS'] = 'tc'
        os.environ['STRAPS_LBS_V'] = '2'
    else:
        os.environ['STRAPS_LBS'] = mode


@pytest.mark.parametrize('B', [1, 7, 32, 40, 64, 127, 200, 700])
def test_tensor_core_lbs_against_oracle_and_cuda_core_kernel(assets_root, smpl_oracle, B):
    """Batches >= 32 take smpl_tc.cu (blend shapes as a 3-pass fp16-split tcgen05 GEMM).  Held to 1e-5 against the oracle
    (smplx lbs()) -- ten times tighter than north_star -- and compared with the CUDA-core kernel on the same
    inputs; ragged body groups (B not a multiple of 64), axis-angle input and transl included."""
    smpl = _smpl(B)
    betas, pose6d, aa = smpl_inputs(B, 900 + B)
    betas *= 2.0                                           # +-6 sigma shapes: large Bm entries
    transl = np.random.RandomState(B).normal(0, 1, (B, 3)).astype(np.float32)
    R, v, j = _oracle_rotmats(smpl_oracle, pose6d, betas)
    Rg, bg, tg = R.to(DEV), torch.from_numpy(betas).to(DEV), torch.from_numpy(transl).to(DEV)
    t = torch.from_numpy(transl)[:, None]
    try:
        outs = {}
        for mode in ('tc', 'tc2', 'simt'):
            _lbs_mode(mode)
            with torch.no_grad():
                o = smpl(body_pose=Rg[:, 1:], global_orient=Rg[:, :1], betas=bg, transl=tg, pose2rot=False)
                oa = smpl(body_pose=torch.from_numpy(aa[:, 3:]).to(DEV), global_orient=torch.from_numpy(aa[:, :3]).to(DEV), betas=bg)
            outs[mode] = (o.vertices.cpu(), o.joints.cpu(), oa.vertices.cpu(), oa.joints.cpu())
    finally:
        _lbs_mode(None)
    with torch.no_grad():
        va, ja = smpl_oracle.forward(betas=torch.from_numpy(betas), body_pose=torch.from_numpy(aa[:, 3:]),
                                     global_orient=torch.from_numpy(aa[:, :3]), pose2rot=True)
    tc, simt = outs['tc'], outs['simt']
    assert rel_err(tc[0].numpy(), (v + t).numpy()) < 1e-5
    assert rel_err(tc[1].numpy(), (j + t).numpy()) < 1e-5
    assert rel_err(tc[2].numpy(), va.numpy()) < 1e-5
    assert rel_err(tc[3].numpy(), ja.numpy()) < 1e-5
    for a, b in zip(tc, simt):
        assert rel_err(a.numpy(), b.numpy()) < 3e-6
    for a, b in zip(tc, outs['tc2']):                       # same MMAs in the same order per output: the two tensor-core kernels agree exactly
        assert torch.equal(a, b)
    assert rel_err(tc[1][:, :24].numpy(), simt[1][:, :24].numpy()) < 1e-6   # chain kernel = the CUDA-core kernel's prologue


def test_tensor_core_lbs_training_outputs(assets_root):
    """The training path saves v_posed and the skinning transforms for the backward kernels: same values from both forward kernels,
    and the gradients that follow from them agree."""
    B = 64
    smpl = _smpl(B)
    betas, pose6d, _ = smpl_inputs(B, 77)
    from utils.rigid_transform_utils import rot6d_to_rotmat
    grads = {}
    try:
        for mode in (None, 'simt'):
            _lbs_mode(mode)
            p = torch.from_numpy(pose6d).to(DEV).requires_grad_(True)
            b = torch.from_numpy(betas).to(DEV).requires_grad_(True)
            R = rot6d_to_rotmat(p).view(B, 24, 3, 3)
            out = smpl(body_pose=R[:, 1:], global_orient=R[:, :1], betas=b, pose2rot=False)
            w = torch.linspace(-1, 1, 6890 * 3, device=DEV).view(1, 6890, 3)
            ((out.vertices * w).sum() + (out.joints ** 2).sum()).backward()
            grads[mode] = (p.grad.cpu(), b.grad.cpu(), out.vertices.detach().cpu())
    finally:
        _lbs_mode(None)
    for a, b in zip(grads[None], grads['simt']):
        assert rel_err(a.numpy(), b.numpy()) < 1e-5


def test_tensor_core_lbs_at_the_sweep_size(assets_root):
    """BASELINE config 5's largest batch (B = 4096: 64 body groups x 54 vertex tiles, 23 work items per persistent CTA): the
    size-independent properties -- identity pose => shape blend, a root rotation is a rigid motion -- and agreement with the CUDA-core
    kernel on random poses."""
    B = 4096
    smpl = _smpl(B)
    rng = np.random.RandomState(12)
    betas = torch.from_numpy(rng.normal(0, 1, (B, 10)).astype(np.float32)).to(DEV)
    R = torch.eye(3, device=DEV).repeat(B, 24, 1, 1)
    from utils.rigid_transform_utils import rot6d_to_rotmat
    with torch.no_grad():
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, :1], betas=betas, pose2rot=False)
        v_shaped = smpl.v_template + torch.einsum('bl,mkl->bmk', betas, smpl.shapedirs)
        J = torch.einsum('bik,ji->bjk', v_shaped, smpl.J_regressor)
        assert rel_err(out.vertices.cpu().numpy(), v_shaped.cpu().numpy()) < 1e-5
        assert rel_err(out.joints[:, :24].cpu().numpy(), J.cpu().numpy()) < 1e-5
        R0 = rot6d_to_rotmat(torch.from_numpy(rng.normal(0, 1, (B, 6)).astype(np.float32)).to(DEV))
        R2 = R.clone()
        R2[:, 0] = R0
        out2 = smpl(body_pose=R2[:, 1:], global_orient=R2[:, :1], betas=betas, pose2rot=False)
        expect = torch.einsum('bij,bvj->bvi', R0, v_shaped - J[:, :1]) + J[:, :1]
        assert rel_err(out2.vertices.cpu().numpy(), expect.cpu().numpy()) < 1e-5
        Rr = rot6d_to_rotmat(torch.from_numpy(rng.normal(0, 1, (B, 144)).astype(np.float32)).to(DEV)).view(B, 24, 3, 3)
        try:
            _lbs_mode('tc')
            a = smpl(body_pose=Rr[:, 1:], global_orient=Rr[:, :1], betas=betas, pose2rot=False)
            _lbs_mode('simt')
            b = smpl(body_pose=Rr[:, 1:], global_orient=Rr[:, :1], betas=betas, pose2rot=False)
        finally:
            _lbs_mode(None)
        assert float((a.vertices - b.vertices).abs().max() / b.vertices.abs().max()) < 3e-6
        assert float((a.joints - b.joints).abs().max() / b.joints.abs().max()) < 3e-6
