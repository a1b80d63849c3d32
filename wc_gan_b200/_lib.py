"""ctypes binding of libwc_b200.so (include/wc_b200.h).  There is no CPU fallback: if the
library is missing it is built with nvcc, and if that fails importing this module raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch  # noqa: F401  (loads the CUDA runtime the library links against)

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

WC_DECOMP = {'cholesky': 0, 'zca': 1}
INFO_SLOTS = 4
INFO_NOT_PD, INFO_BAD_CLASS, INFO_JACOBI, INFO_XCHG_TIMEOUT = 0, 1, 2, 3


class WcError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    # build() is a no-op when the source digest matches the stamp next to the .so, so a library left behind by an
    # older csrc/ or header is never loaded silently; it serialises concurrent ranks with a file lock.
    path = _build.build(force=bool(os.environ.get('WC_B200_REBUILD')))
    try:
        return ctypes.CDLL(path)
    except OSError as e:  # fail loudly: the product has no other implementation
        raise WcError(f'cannot load {path}: {e}') from e


lib = _load()


class WcLayerCall(ctypes.Structure):
    """include/wc_b200.h: struct WcLayerCall (field order and types must match)."""
    _fields_ = ([(n, c_int32) for n in ('N', 'HW', 'Cin', 'Cout', 'g', 'K', 'E', 'whiten', 'decomposition', 'renorm',
                                        'training', 'triangular', 'phases', 'use_tensor_cores', 'relu', 'reserved0')] +
                [('eps', c_float), ('momentum', c_float), ('count', c_double)] +
                [(n, c_void_p) for n in ('count_dev', 'Wu', 'bu', 'Wc', 'bc', 'class_matrix', 'D', 'bf', 'moving_mean',
                                         'moving_cov', 'cls', 'info', 'x', 'y', 'mean_ext', 'S_ext', 'dy', 'dx', 'dWu',
                                         'dbu', 'dWc', 'dbc', 'dclass_matrix', 'dD', 'dbf', 'red')] +
                [('saved', c_void_p), ('saved_bytes', c_size_t), ('workspace', c_void_p), ('workspace_bytes', c_size_t)] +
                [('xchg_peers', c_void_p), ('xchg_state', c_void_p), ('xchg_rank', c_int32), ('xchg_world', c_int32)])


class WcSavedLayout(ctypes.Structure):
    _fields_ = [(n, c_size_t) for n in ('rec', 'mean', 'S', 'rec_count', 'Weff', 'Wm', 'R', 'U', 'lam', 'Wk', 'T', 'b',
                                        'ehat', 'Tp_fwd', 'Tp_bwd', 'A', 'cvec', 'total')]


_P = c_void_p
_SIGS = {
    'wc_version': (c_int, []),
    'wc_sizeof_layer_call': (c_size_t, []),
    'wc_sizeof_saved_layout': (c_size_t, []),
    'wc_launch_count': (ctypes.c_longlong, []),
    'wc_path_for': (c_int, [c_int64, c_int, c_int, c_int]),
    'wc_set_tensor_core_gram': (None, [c_int]),
    'wc_stats_workspace_bytes': (c_size_t, [c_int64, c_int, c_int]),
    'wc_stats_fwd': (c_int, [_P, c_int64, c_int, c_int, _P, c_size_t, _P, _P, _P]),
    'wc_stats_merge': (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P]),
    'wc_factor_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'wc_factor_fwd': (c_int, [_P, _P, _P, c_double, c_int, c_int, c_float, c_float, c_int, c_int, c_int,
                              _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    'wc_compose_fwd': (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P, c_int, _P,
                               _P, _P, _P, _P, _P]),
    'wc_apply_fwd': (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P]),
    'wc_pack_bytes': (c_size_t, [c_int, c_int, c_int]),
    'wc_pack_fwd': (c_int, [_P, c_int, c_int, _P, _P]),
    'wc_pack_bwd': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    'wc_reduce_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    'wc_reduce_bwd': (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P, c_size_t, _P]),
    'wc_color_bwd_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    'wc_color_bwd': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int,
                             _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    'wc_factor_bwd_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'wc_factor_bwd': (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_double, c_int, c_int, c_float, c_int, c_int,
                              _P, _P, _P, c_size_t, _P]),
    'wc_apply_bwd': (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P, c_int, _P, _P]),
    'wc_exchange_bytes': (c_size_t, [c_int, c_int, c_int]),
    'wc_layer_saved_layout': (c_int, [ctypes.POINTER(WcLayerCall), ctypes.POINTER(WcSavedLayout)]),
    'wc_layer_saved_bytes': (c_size_t, [ctypes.POINTER(WcLayerCall)]),
    'wc_layer_workspace_bytes': (c_size_t, [ctypes.POINTER(WcLayerCall), c_int]),
    'wc_layer_fwd': (c_int, [ctypes.POINTER(WcLayerCall), _P]),
    'wc_layer_bwd': (c_int, [ctypes.POINTER(WcLayerCall), _P]),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTS = tuple(_SIGS.keys())

# the ctypes mirrors above must describe the structs the library was compiled with
if lib.wc_version() != 102:
    raise WcError(f'libwc_b200.so reports version {lib.wc_version()}, this package expects 102 (stale build?)')
if (lib.wc_sizeof_layer_call() != ctypes.sizeof(WcLayerCall) or
        lib.wc_sizeof_saved_layout() != ctypes.sizeof(WcSavedLayout)):
    raise WcError('WcLayerCall / WcSavedLayout layout mismatch between include/wc_b200.h and wc_gan_b200/_lib.py')

_ERRS = {-1: 'invalid argument', -2: 'workspace too small', -3: 'unsupported shape', -4: 'misaligned pointer'}


def check(rc: int, what: str) -> None:
    """Negative codes are argument errors (the reference raises ValueError / assert there,
    gan/conditional_layers.py:615,628-632,768-770); positive codes are cudaError_t."""
    if rc == 0:
        return
    if rc < 0:
        raise ValueError(f'{what}: {_ERRS.get(rc, rc)}')
    raise WcError(f'{what}: CUDA error {rc}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
