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


def _run_c_client(tmp_path):
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "openfermion-fqe_b200", "fqe_b200", "lib")
    exe = str(tmp_path / "c_client")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "examples", "c_client.c"), "-o", exe,
                           "-L" + libdir, "-lfqe_b200", "-Wl,-rpath," + libdir, "-lm"])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "libfqe_b200 version" in res.stdout
    return res.stdout


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_plain_c_client_links_and_reports_missing_device(tmp_path):
    """examples/c_client.c: the boundary is usable from plain C (no Python / torch / C++ in
    the signatures).  It must compile against include/fqe_b200.h, link to the library and - in
    a container without a GPU - get FQEB_ERR_NODEVICE instead of a CPU fallback."""
    out = _run_c_client(tmp_path)
    assert "no device" in out and "no CPU fallback" in out


@pytest.mark.gpu
def test_plain_c_client_builds_sigma_on_the_gpu(tmp_path):
    """the same client on a GPU box: graph, operator and sigma through the C ABI alone"""
    out = _run_c_client(tmp_path)
    assert "|sigma| =" in out and "no device" not in out
