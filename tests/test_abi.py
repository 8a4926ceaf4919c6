"""The C-ABI library loads on a CPU-only box and exports every symbol include/efa_xray_b200.h declares.
No compute calls are made here."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, 'include', 'efa_xray_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(exb_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'library does not export %s' % n


def test_binding_covers_header():
    from efa_xray_b200 import _lib
    assert set(_declared()) == set(_lib.SIGNATURES) | {'exb_last_error', 'exb_launch_count',
                                                        'exb_state_sweep_row_granularity'}


def test_version_and_loud_failure_without_device(lib):
    import torch
    from efa_xray_b200 import _lib
    assert lib.exb_version() == 100
    if not torch.cuda.is_available():
        with pytest.raises(_lib.ExbError):
            _lib.require_device()
        with pytest.raises(_lib.ExbError):
            _lib.call('exb_device_check')


def test_argument_validation_returns_status(lib):
    # null pointers are rejected before any CUDA call
    assert lib.exb_grid_unitvec(None, None, 0, None, None) == -1
    assert b'exb_grid_unitvec' in lib.exb_last_error()


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (no CPU fallback)."""
    bad = []
    for pkg in ('efa_xray_b200', 'efa_xray'):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh', '.h')):
                    src = open(os.path.join(dirpath, f)).read()
                    if re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
