"""The C-ABI library loads and exports every symbol include/fqe_b200.h declares.
No compute call needs a GPU here; without one the entry points must fail loudly
(FQEB_ERR_NODEVICE), never fall back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from fqe_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fqe_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fqeb_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    decl = declared_symbols()
    assert len(decl) >= 30
    assert sorted(L.SIGNATURES) == decl


def test_every_symbol_is_exported():
    lib = L.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_scalars():
    lib = L.load()
    assert lib.fqeb_version() >= 100
    assert lib.fqeb_gemm_col_align() == 128
    assert lib.fqeb_reduce_scratch_bytes() > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    lib = L.load()
    assert lib.fqeb_device_count() == 0
    h = ctypes.c_void_p()
    rc = lib.fqeb_graph_create(4, 2, 2, ctypes.byref(h))
    assert rc == L.ERR_NODEVICE
    assert b"no CPU fallback" in lib.fqeb_last_error()
    x = np.zeros(4, dtype=np.complex128)
    assert lib.fqeb_zscal(4, 1.0, 0.0, x.ctypes.data, None) == L.ERR_NODEVICE
    import fqe_b200
    with pytest.raises(L.FqeB200Error):
        fqe_b200.get_wavefunction(2, 0, 2)
