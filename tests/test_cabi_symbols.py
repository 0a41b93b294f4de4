"""The C-ABI library loads and exports exactly what include/sgnn_b200.h declares (no compute calls)."""
import os
import re
import subprocess

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, 'include', 'sgnn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sgnn_[A-Za-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound():
    from sgnn_b200 import _lib
    names = _header_functions()
    assert len(names) >= 20
    assert sorted(_lib.SIGNATURES) == names
    out = subprocess.check_output(['nm', '-D', '--defined-only', _lib.LIB_PATH]).decode()
    exported = set(re.findall(r' T (sgnn_[A-Za-z0-9_]+)', out))
    assert set(names) <= exported
    for n in names:
        assert getattr(_lib.lib, n) is not None


def test_host_only_entry_points():
    from sgnn_b200 import _lib
    assert _lib.lib.sgnn_version() == 100
    assert _lib.lib.sgnn_error_string(0) == b'ok'
    assert b'invalid' in _lib.lib.sgnn_error_string(-1)
    assert _lib.lib.sgnn_scan_scratch_bytes(10 ** 7) > 4 * (10 ** 7) // 2048
    assert _lib.lib.sgnn_compact_scratch_bytes(1000) > 1000 + 4 * 1001


def test_library_is_sm100a_and_uses_no_cpu_fallback():
    from sgnn_b200 import _lib
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    import sgnn_b200.engine as E
    import torch
    import pytest
    with pytest.raises(RuntimeError):
        E.coords_to_i64(torch.zeros((4, 4), dtype=torch.int32))      # CPU tensor -> loud failure


def test_export_argument_validation_without_a_launch():
    """sgnn_export rejects malformed segment lists on the host (no kernel is launched for these)."""
    import ctypes as C
    from sgnn_b200 import _lib
    segs = (_lib.SgnnExportSeg * 17)()
    assert _lib.lib.sgnn_export(segs, 0, None) == 0                      # nothing to do
    assert _lib.lib.sgnn_export(segs, 17, None) != 0                     # more than SGNN_EXPORT_MAX_SEGS
    assert _lib.lib.sgnn_export(None, 1, None) != 0
    segs[0].n, segs[0].kind = 5, 7                                       # unknown kind
    assert _lib.lib.sgnn_export(segs, 1, None) != 0
    segs[0].kind, segs[0].dst = _lib.EXPORT_COPY32, None                 # non-empty segment without a destination
    assert _lib.lib.sgnn_export(segs, 1, None) != 0
    segs[0].n = -1
    assert _lib.lib.sgnn_export(segs, 1, None) != 0
