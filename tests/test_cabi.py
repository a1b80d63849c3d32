"""C-ABI surface checks that need no GPU: the library builds/loads, exports every symbol that
include/wc_b200.h declares, and rejects bad arguments before touching the device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'wc_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(wc_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    from wc_gan_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(_lib.lib, n), f'{n} declared in include/wc_b200.h but not exported'
    assert set(names) == set(_lib.EXPORTS), 'ctypes signatures and header disagree'
    assert _lib.lib.wc_version() == 102
    # the diagnostic entry points are declared too (include/wc_b200_debug.h), and nothing else is exported
    dbg = re.sub(r'/\*.*?\*/', '', open(os.path.join(ROOT, 'include', 'wc_b200_debug.h')).read(), flags=re.S)
    dbg_names = sorted(set(re.findall(r'\b(wc_[a-z0-9_]+)\s*\(', dbg)))
    for n in dbg_names:
        assert hasattr(_lib.lib, n), n
    import subprocess
    out = subprocess.run(['nm', '-D', '--defined-only', _lib._build.LIB], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ' T ' in ln and ln.split()[-1].startswith('wc_')}
    assert exported == set(names) | set(dbg_names), exported ^ (set(names) | set(dbg_names))


def test_argument_errors_without_device():
    from wc_gan_b200 import _lib
    L = _lib.lib
    assert L.wc_stats_fwd(None, 16, 8, 8, None, 0, None, None, None) == -1
    assert L.wc_stats_workspace_bytes(16, 8, 3) == 0          # C % g != 0
    assert L.wc_stats_workspace_bytes(65536, 256, 256) > 0
    assert L.wc_apply_fwd(None, 1, 1, 1, 1, None, None, None, None, None, 1, 0, None, None, None) == -1
    assert L.wc_pack_bytes(256, 1, 0) == 256 * 256 * 4 and L.wc_pack_bytes(100, 1, 0) == 0
    assert L.wc_path_for(65536, 256, 256, 256) & 1 and L.wc_path_for(100, 24, 24, 24) == 0
    assert L.wc_reduce_bwd(None, None, 1, 1, 1, 1, None, None, 1, None, None, None, 0, None) == -1
    assert L.wc_factor_fwd(None, None, None, 1.0, 8, 3, 1e-3, 0.99, 0, 0, 1, None, None, None, None, None, None, None,
                           None, None, 0, None) == -1
    # peer-memory exchange buffer: flags + 2 slots x world x (record + adjoint); 1..16 ranks
    n1, n8 = L.wc_exchange_bytes(256, 256, 1), L.wc_exchange_bytes(256, 256, 8)
    assert n1 >= 1024 + 2 * 4 * (2 * 256 * 256 + 2 * 256 + 1) and n8 > 7 * (n1 - 1024)
    assert L.wc_exchange_bytes(256, 256, 17) == 0 and L.wc_exchange_bytes(256, 7, 2) == 0
    assert ctypes.sizeof(_lib.WcLayerCall) % 8 == 0
    with pytest.raises(ValueError):
        _lib.check(-1, 'x')
    with pytest.raises(_lib.WcError):
        _lib.check(700, 'x')


def test_layers_have_no_cpu_path():
    import torch
    import wc_gan_b200 as W
    for layer, inp in ((W.DecorelationNormalization(), torch.zeros(2, 2, 2, 4)),
                       (W.Conv11(filters=4), torch.zeros(2, 2, 2, 4)),
                       (W.ConditionalConv11(filters=4, number_of_classes=2), [torch.zeros(2, 2, 2, 4), torch.zeros(2, 1, dtype=torch.int32)])):
        with pytest.raises(RuntimeError):
            layer(inp)


def test_create_norm_vocabulary():
    import wc_gan_b200 as W
    for norm in ('n', 'd', 'dr'):
        for an in ('uconv', 'cconv', 'ucconv', 'fconv', 'ufconv', 'n'):
            s = W.create_norm(norm, an, number_of_classes=3, filters_emb=2)(axis=-1, name='G.0_bn1')
            s.build(8, None)
            keys = set(s.state_dict().keys())
            if norm != 'n':
                assert {'npart.moving_mean', 'npart.moving_variance'} <= keys
                assert tuple(s.npart.moving_mean.shape) == (8, 1) and tuple(s.npart.moving_variance.shape) == (8, 8)
            if an == 'ucconv':
                assert tuple(s.repart_c.kernel.shape) == (3, 1, 1, 8, 8) and tuple(s.repart_u.kernel.shape) == (1, 1, 8, 8)
            if an == 'ufconv':
                assert tuple(s.repart_c.class_matrix.shape) == (3, 2) and tuple(s.repart_c.kernel.shape) == (2, 1, 1, 8, 8)
                assert s.repart_c.bias is None            # generator.py:74 use_bias=False
    with pytest.raises(AssertionError):
        W.create_norm('x', 'uconv')
    with pytest.raises(NotImplementedError):
        W.create_norm('d', 'ccs')
