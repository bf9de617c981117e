"""CPU: the C-ABI library loads without a GPU and exports every symbol include/straps_b200.h declares."""
import ctypes
import os
import re

from conftest import REPO
from straps_b200 import _lib


def _declared():
    src = open(os.path.join(REPO, 'include', 'straps_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(straps_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 18
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_version_and_error_string():
    L = _lib.lib()
    assert L.straps_abi_version() == 5
    assert isinstance(L.straps_last_error(), bytes)
    assert L.straps_launch_count() >= 0
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert L.straps_rot6d_to_rotmat(None, 4, None, None) != 0
    assert b'null' in L.straps_last_error()
    h = ctypes.c_void_p()
    assert L.straps_regressor_create(ctypes.byref(h), 99, 4) != 0
    assert b'c_in' in L.straps_last_error()


def test_shipped_kernels_carry_tcgen05_and_tma_sass():
    """The built library's shipped convolution kernels really are tcgen05 / TMEM / TMA code (B200_PROFILING.md mnemonics):
    UTCHMMA = tcgen05.mma, UTMALDG = TMA tile load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit; the LBS and IEF kernels stage their
    constants with bulk-async copies (UBLKCP) on mbarriers (SYNCS).  Checked on the .so the tests load (cuobjdump, no GPU needed)."""
    import shutil
    import sys
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not on PATH')
    sys.path.insert(0, os.path.join(REPO, 'tools'))
    import sass_summary
    from straps_b200 import _lib
    counts = sass_summary.summary(_lib.LIB_PATH)
    names = sass_summary.demangle(list(counts))
    by_name = {names[k].replace('(int)', '').replace('(bool)', '').split('(')[0].replace('void ', ''): v for k, v in counts.items()}
    for shipped in ('straps::conv_tc_kernel<64>', 'straps::conv_tc_kernel<128>', 'straps::conv1_s2d_kernel<1>', 'straps::conv1_s2d_kernel<0>',
                    'straps::wgrad_tc_kernel<64>', 'straps::wgrad_tc_kernel<128>'):
        assert shipped in by_name, (shipped, sorted(k for k in by_name if 'conv' in k or 'wgrad' in k))
        c = by_name[shipped]
        assert c['UTCHMMA'] >= 8 and c['UTMALDG'] >= 3 and c['LDTM'] >= 2 and c['UTCBAR'] >= 2 and c['SYNCS'] >= 4, (shipped, dict(c))
    assert any(k.startswith('straps::lbs_kernel<') and v['UBLKCP'] > 0 for k, v in by_name.items())
    assert by_name['straps::ief_kernel']['UBLKCP'] > 0
