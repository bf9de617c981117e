"""Host-side binding of libstraps_b200.so (hand-written sm_100a CUDA) for the STRAPS hot path.

The directory that contains this package (`straps-3dhumanshapepose_b200/`) is a drop-in source root:
put it on sys.path where the reference repo root used to be and `models.regressor`,
`models.smpl_official`, `utils.*`, `losses.*`, `data.*`, `config` resolve to the B200 path.
There is deliberately no CPU fallback: every op raises if the CUDA library or a CUDA tensor is missing.
"""
from . import _lib  # noqa: F401
