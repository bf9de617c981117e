"""Seeded synthetic SMPL-shaped assets.

The real SMPL files are licence-gated downloads (reference README.md:37-59) and are not
shipped with the reference (additional/README.md:1).  Benchmarks, the smoke test and the parity
tests therefore run on synthetic assets that have *exactly* the keys / shapes / dtypes of the real
files, written in the directory layout `config.py` expects (reference config.py:3-10):

    <root>/additional/smpl/SMPL_NEUTRAL.pkl
    <root>/additional/J_regressor_extra.npy          [9, 6890]
    <root>/additional/cocoplus_regressor.npy         [19, 6890]
    <root>/additional/J_regressor_h36m.npy           [17, 6890]
    <root>/additional/neutral_smpl_mean_params_6dpose.npz   keys pose[144], shape[10]
    <root>/additional/smpl_faces.npy                 [13776, 3]

Values are scaled like the real model (template ~ metres, shapedirs ~1e-2, posedirs ~1e-3) so
that relative-error thresholds mean something.  Everything is drawn from numpy RandomState so the
files are bit-identical on every machine.  A user who owns the real files just drops them in place
of these; no code path depends on the data being synthetic.
"""
import os
import pickle

import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_FACES = 13776
# kintree_table[0] of the SMPL model (parent of each joint; root = 2**32-1 on disk)
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def _rest_skeleton(rng):
    """A rough humanoid skeleton in metres (only needs to be non-degenerate)."""
    J = np.zeros((NUM_JOINTS, 3))
    step = {0: (0, -0.22, 0.0), 1: (0.07, -0.09, 0), 2: (-0.07, -0.09, 0), 3: (0, 0.11, -0.02),
            4: (0.03, -0.38, 0), 5: (-0.03, -0.38, 0), 6: (0, 0.13, 0.02), 7: (-0.01, -0.40, -0.03),
            8: (0.01, -0.40, -0.03), 9: (0, 0.06, 0.02), 10: (0.03, -0.06, 0.12), 11: (-0.03, -0.06, 0.12),
            12: (0, 0.21, -0.04), 13: (0.08, 0.11, -0.02), 14: (-0.08, 0.11, -0.02), 15: (0, 0.09, 0.05),
            16: (0.11, 0.04, -0.01), 17: (-0.11, 0.04, -0.01), 18: (0.26, -0.01, -0.02),
            19: (-0.26, -0.01, -0.02), 20: (0.25, 0.0, 0.0), 21: (-0.25, 0.0, 0.0),
            22: (0.08, -0.01, -0.01), 23: (-0.08, -0.01, -0.01)}
    for j in range(NUM_JOINTS):
        p = SMPL_PARENTS[j]
        base = J[p] if p >= 0 else np.zeros(3)
        J[j] = base + np.asarray(step[j]) + rng.normal(0, 0.004, 3)
    return J


def _sparse_convex_rows(rng, n_rows, centres, verts, nnz_lo, nnz_hi):
    """Rows that are convex combinations of the vertices nearest to `centres[r]`."""
    R = np.zeros((n_rows, NUM_VERTS))
    for r in range(n_rows):
        d = np.linalg.norm(verts - centres[r][None], axis=1)
        k = int(rng.randint(nnz_lo, nnz_hi + 1))
        idx = np.argsort(d, kind='stable')[:k]
        w = rng.gamma(1.0, 1.0, k) + 1e-3
        R[r, idx] = w / w.sum()
    return R


def build_smpl_dict(seed=0):
    rng = np.random.RandomState(seed)
    J = _rest_skeleton(rng)
    # every vertex hangs off a primary joint
    primary = rng.randint(0, NUM_JOINTS, NUM_VERTS)
    primary[:NUM_JOINTS * 8] = np.repeat(np.arange(NUM_JOINTS), 8)   # every joint owns vertices
    v_template = J[primary] + rng.normal(0, 0.045, (NUM_VERTS, 3))

    children = [[c for c in range(NUM_JOINTS) if SMPL_PARENTS[c] == j] for j in range(NUM_JOINTS)]
    weights = np.zeros((NUM_VERTS, NUM_JOINTS))
    for v in range(NUM_VERTS):
        j = primary[v]
        near = [j] + ([SMPL_PARENTS[j]] if SMPL_PARENTS[j] >= 0 else []) + children[j]
        k = int(rng.randint(1, 5))                       # 1..4 non-zeros per vertex, like real SMPL
        pick = near[:k] if len(near) >= k else near
        w = rng.gamma(2.0, 1.0, len(pick))
        w[0] += 1.0
        weights[v, pick] = w / w.sum()

    shapedirs = rng.normal(0, 1.0e-2, (NUM_VERTS, 3, NUM_BETAS))
    posedirs = rng.normal(0, 1.0e-3, (NUM_VERTS, 3, (NUM_JOINTS - 1) * 9))
    J_regressor = _sparse_convex_rows(rng, NUM_JOINTS, J, v_template, 12, 40)

    kintree = np.zeros((2, NUM_JOINTS), dtype=np.uint32)
    kintree[0] = np.asarray([p if p >= 0 else 2 ** 32 - 1 for p in SMPL_PARENTS], dtype=np.uint32)
    kintree[1] = np.arange(NUM_JOINTS, dtype=np.uint32)

    faces = np.zeros((NUM_FACES, 3), dtype=np.uint32)
    a = rng.randint(0, NUM_VERTS, NUM_FACES)
    faces[:, 0] = a
    faces[:, 1] = (a + rng.randint(1, 40, NUM_FACES)) % NUM_VERTS
    faces[:, 2] = (a + rng.randint(40, 90, NUM_FACES)) % NUM_VERTS

    import scipy.sparse
    return {
        'v_template': v_template.astype(np.float64),
        'shapedirs': shapedirs.astype(np.float64),
        'posedirs': posedirs.astype(np.float64),
        'J_regressor': scipy.sparse.csc_matrix(J_regressor),
        'weights': weights.astype(np.float64),
        'kintree_table': kintree,
        'f': faces,
    }, J, v_template


def write_synthetic_assets(root, seed=0, overwrite=False):
    """Write all asset files below `<root>/additional/`.  Returns the `additional` dir."""
    add = os.path.join(root, 'additional')
    smpl_dir = os.path.join(add, 'smpl')
    done = os.path.join(add, '.synthetic_seed_%d' % seed)
    if os.path.exists(done) and not overwrite:
        return add
    os.makedirs(smpl_dir, exist_ok=True)
    smpl, J, v_template = build_smpl_dict(seed)
    with open(os.path.join(smpl_dir, 'SMPL_NEUTRAL.pkl'), 'wb') as f:
        pickle.dump(smpl, f, protocol=2)

    rng = np.random.RandomState(seed + 1)

    def centres(n):
        return J[rng.randint(0, NUM_JOINTS, n)] + rng.normal(0, 0.03, (n, 3))

    np.save(os.path.join(add, 'J_regressor_extra.npy'),
            _sparse_convex_rows(rng, 9, centres(9), v_template, 4, 24).astype(np.float32))
    np.save(os.path.join(add, 'cocoplus_regressor.npy'),
            _sparse_convex_rows(rng, 19, centres(19), v_template, 8, 48).astype(np.float32))
    np.save(os.path.join(add, 'J_regressor_h36m.npy'),
            _sparse_convex_rows(rng, 17, centres(17), v_template, 16, 96).astype(np.float32))
    np.save(os.path.join(add, 'smpl_faces.npy'), smpl['f'].astype(np.int64))

    # mean parameters: near-identity 6-D rotations (interleaved a1/a2 layout, see rot6d_to_rotmat)
    ident6 = np.array([1., 0., 0., 1., 0., 0.])
    pose = np.tile(ident6, NUM_JOINTS) + rng.normal(0, 0.08, NUM_JOINTS * 6)
    shape = rng.normal(0, 0.3, NUM_BETAS)
    np.savez(os.path.join(add, 'neutral_smpl_mean_params_6dpose.npz'), pose=pose, shape=shape)
    with open(done, 'w') as f:
        f.write('synthetic assets, seed %d\n' % seed)
    return add


def write_synthetic_dataset(npz_path, n=32, seed=0):
    """A SyntheticTrainingDataset-shaped npz (reference data/synthetic_training_dataset.py:21-24)."""
    rng = np.random.RandomState(seed)
    prefixes = ['h36m', 'up3d', '3dpw', 'amass']
    fnames = np.array(['%s_%05d' % (prefixes[i % 4], i) for i in range(n)])
    poses = rng.normal(0, 0.25, (n, 72)).astype(np.float64)
    shapes = rng.normal(0, 1.0, (n, 10)).astype(np.float64)
    np.savez(npz_path, fnames=fnames, poses=poses, shapes=shapes)
    return npz_path
