"""ORACLE / TEST INFRASTRUCTURE -- load the UNMODIFIED reference modules from /root/reference.

Only usable in the authoring container (the GPU box has no /root/reference); used to pin
oracle/straps_oracle.py and to generate tests/golden/*.npz (oracle/gen_golden.py).

The reference imports `config`, `models.*`, `utils.*`, `losses.*` as top-level names relative to
its repo root and resolves asset paths relative to the CWD (config.py:3-10), and needs the absent
third-party `smplx` (-> oracle/smplx_shim).  Our product tree uses the very same top-level names
(it is a drop-in), so the reference modules are imported under a temporary sys.path / CWD and then
moved out of sys.modules into a private namespace.
"""
import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('STRAPS_REFERENCE_ROOT', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, 'smplx_shim')
_TOP = ('config', 'models', 'utils', 'losses', 'data', 'augmentation', 'metrics')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'models'))


@contextlib.contextmanager
def _isolated(assets_root):
    saved_modules = {k: v for k, v in sys.modules.items() if k.split('.')[0] in _TOP}
    for k in saved_modules:
        del sys.modules[k]
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    sys.path[:] = [REFERENCE_ROOT, _SHIM] + [p for p in saved_path if 'straps-3dhumanshapepose_b200' not in p]
    os.chdir(assets_root)
    try:
        yield
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split('.')[0] in _TOP]:
            del sys.modules[k]
        sys.modules.update(saved_modules)


def load_reference(assets_root):
    """Returns a namespace with the reference's own classes/functions for the hot path.

    `assets_root` must contain `additional/...` (see straps_b200.synthetic_assets).
    """
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    with _isolated(assets_root):
        ns.config = importlib.import_module('config')
        ns.SingleInputRegressor = importlib.import_module('models.regressor').SingleInputRegressor
        ns.SMPL = importlib.import_module('models.smpl_official').SMPL
        ns.rot6d_to_rotmat = importlib.import_module('utils.rigid_transform_utils').rot6d_to_rotmat
        ns.orthographic_project_torch = importlib.import_module('utils.cam_utils').orthographic_project_torch
        ns.check_joints2d_visibility_torch = importlib.import_module('utils.joints2d_utils').check_joints2d_visibility_torch
        ns.Loss = importlib.import_module('losses.multi_task_loss').HomoscedasticUncertaintyWeightedMultiTaskLoss
        ns.SyntheticTrainingDataset = importlib.import_module('data.synthetic_training_dataset').SyntheticTrainingDataset
        ns.heatmaps = importlib.import_module('utils.label_conversions').convert_2Djoints_to_gaussian_heatmaps_torch
        # SURVEY 8f rows N2 / N4
        ns.smpl_augmentation = importlib.import_module('augmentation.smpl_augmentation')
        ns.cam_augmentation = importlib.import_module('augmentation.cam_augmentation')
        ns.perspective_project_torch = importlib.import_module('utils.cam_utils').perspective_project_torch
        ns.get_intrinsics_matrix = importlib.import_module('utils.cam_utils').get_intrinsics_matrix
        ns.eval_utils = importlib.import_module('utils.eval_utils')
        ns.Tracker = importlib.import_module('metrics.train_loss_and_metrics_tracker').TrainingLossesAndMetricsTracker
    ns.assets_root = assets_root

    @contextlib.contextmanager
    def in_assets_cwd():
        cwd = os.getcwd()
        os.chdir(assets_root)
        try:
            yield
        finally:
            os.chdir(cwd)
    ns.cwd = in_assets_cwd
    return ns
