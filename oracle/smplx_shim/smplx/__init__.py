"""ORACLE / TEST INFRASTRUCTURE -- not shipped, never imported by the product path.

Minimal CPU restatement of the third-party `smplx` package (PyPI, un-pinned in the reference's
requirements.txt:6; effective pin <= 0.1.21 because models/smpl_official.py:4 imports
`smplx.body_models.ModelOutput`).  The package is absent from /root/reference and cannot be
installed (no network), so the published algorithm (SMPL, Loper et al. 2015; smplx/lbs.py) is
restated here from SURVEY.md section 8a rows S1-S8 / Appendix A.  PARITY UNPINNED for this part:
the reference holds no golden vectors for SMPL outputs; the restatement is anchored on the
reference's own call sites (models/smpl_official.py:15-41, train/...:132-206) which run unchanged
on top of it, and on the mathematical invariants tested in tests/test_oracle_invariants.py.
"""
from .body_models import SMPL, ModelOutput, create  # noqa: F401
from . import lbs  # noqa: F401
