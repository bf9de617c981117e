"""Constants of the STRAPS hot path, value-for-value those of the reference's config.py:1-34.

Asset paths are relative to the current working directory, like the reference's; set the
environment variable STRAPS_ASSETS_ROOT to resolve them against another directory instead.
"""
import os as _os

_ROOT = _os.environ.get('STRAPS_ASSETS_ROOT', '')


def _asset(rel):
    return _os.path.join(_ROOT, rel) if _ROOT else rel


# --- files (reference config.py:3-10)
SMPL_MODEL_DIR = _asset('additional/smpl')
SMPL_FACES_PATH = _asset('additional/smpl_faces.npy')
SMPL_MEAN_PARAMS_PATH = _asset('additional/neutral_smpl_mean_params_6dpose.npz')
J_REGRESSOR_EXTRA_PATH = _asset('additional/J_regressor_extra.npy')
COCOPLUS_REGRESSOR_PATH = _asset('additional/cocoplus_regressor.npy')
H36M_REGRESSOR_PATH = _asset('additional/J_regressor_h36m.npy')
VERTEX_TEXTURE_PATH = _asset('additional/vertex_texture.npy')
CUBE_PARTS_PATH = _asset('additional/cube_parts.npy')

# --- scalars (reference config.py:13-14)
FOCAL_LENGTH = 5000.
REGRESSOR_IMG_WH = 256

# --- joint conventions (reference config.py:27-32).  The SMPL module returns 90 joints:
# 24 posed + 21 vertex picks, then 9 "extra", 19 cocoplus and 17 H36M regressed joints.
ALL_JOINTS_TO_COCO_MAP = [24, 26, 25, 28, 27] + [16, 17, 18, 19, 20, 21] + [1, 2, 4, 5, 7, 8]
ALL_JOINTS_TO_H36M_MAP = [73 + i for i in range(17)]
H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]
H36M_TO_J14 = H36M_TO_J17[:14]
