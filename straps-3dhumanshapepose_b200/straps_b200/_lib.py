"""ctypes loader for libstraps_b200.so (C ABI declared in include/straps_b200.h)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# STRAPS_B200_LIB selects another build of the same ABI (tools/conv_experiment.sh uses the -DSTRAPS_TC_EXPERIMENTS one)
LIB_PATH = os.environ.get('STRAPS_B200_LIB') or os.path.join(_HERE, 'libstraps_b200.so')
CSRC_DIR = os.path.join(os.path.dirname(_HERE), 'csrc')

c_float_p = ctypes.POINTER(ctypes.c_float)
c_i64_p = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/straps_b200.h one to one
SIGNATURES = {
    'straps_last_error': (ctypes.c_char_p, []),
    'straps_abi_version': (ctypes.c_int, []),
    'straps_launch_count': (ctypes.c_ulonglong, []),
    'straps_smpl_create': (ctypes.c_int, [ctypes.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'straps_smpl_destroy': (None, [_vp]),
    'straps_smpl_is_sparse4': (ctypes.c_int, [_vp]),
    'straps_smpl_forward': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, ctypes.c_int64, _vp, ctypes.c_int64, _vp,
                                           ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    'straps_smpl_forward_train': (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    'straps_smpl_backward': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp]),
    'straps_rot6d_backward': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, _vp]),
    'straps_orthographic_project_backward': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    'straps_rot6d_to_rotmat': (ctypes.c_int, [_vp, ctypes.c_int64, _vp, _vp]),
    'straps_orthographic_project': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    'straps_joints2d_to_heatmaps': (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    'straps_multiclass_to_binary': (ctypes.c_int, [_vp, ctypes.c_int64, _vp, _vp]),
    'straps_batch_rodrigues': (ctypes.c_int, [_vp, ctypes.c_int64, _vp, _vp]),
    'straps_batch_rodrigues_backward': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, _vp]),
    'straps_perspective_project': (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    'straps_scale_shift': (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    'straps_points_metrics': (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp]),
    'straps_rows_metric': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, _vp, _vp]),
    'straps_accumulate': (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double, _vp, _vp]),
    'straps_regressor_create': (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int]),
    'straps_regressor_destroy': (None, [_vp]),
    'straps_regressor_workspace_bytes': (ctypes.c_size_t, [_vp]),
    'straps_regressor_conv_name': (ctypes.c_char_p, [_vp, ctypes.c_int]),
    'straps_regressor_load': (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                                             ctypes.POINTER(_vp), _vp, _vp]),
    'straps_encoder_forward': (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    'straps_ief_forward': (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    'straps_regressor_forward': (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    'straps_regressor_forward_from_labels': (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                           _vp, _vp, _vp]),
    'straps_encoder_train_forward': (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    'straps_encoder_backward': (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp]),
    'straps_encoder_backward_range': (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    'straps_ief_forward_train': (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    'straps_ief_backward': (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.POINTER(_vp),
                                           ctypes.POINTER(_vp), _vp, _vp]),
    'straps_adam_step': (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                        ctypes.c_float, ctypes.c_float, ctypes.c_float, _vp]),
    'straps_adam_step_dev': (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, _vp, ctypes.c_float, ctypes.c_float,
                                            ctypes.c_float, ctypes.c_float, ctypes.c_float, _vp]),
    'straps_multitask_loss': (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), c_i64_p, _vp, ctypes.c_int,
                                             ctypes.c_float, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    'straps_encoder_read_activation': (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_int, _vp, c_i64_p, _vp]),
}

_lib = None


class StrapsError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libstraps_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(['make', '-C', CSRC_DIR, '-j8'], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise StrapsError('building libstraps_b200.so failed (see output above)')
    return LIB_PATH


def lib():
    """The loaded library; fails loudly (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StrapsError('%s is missing -- run `python -c "import __graft_entry__ as g; g.build()"` or '
                              '`make -C %s`; there is no CPU fallback' % (LIB_PATH, CSRC_DIR))
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().straps_last_error()
        raise StrapsError('%s failed (rc=%d): %s' % (what or 'libstraps_b200 call', rc, msg.decode() if msg else '?'))


def launch_count():
    return int(lib().straps_launch_count())
