"""pytest configuration: registers the ``gpu`` marker and makes the repo root
(for ``oracle``) and ``openfermion-fqe_b200`` (for ``fqe_b200``) importable."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "openfermion-fqe_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# build the C-ABI library if it is missing (nvcc cross-compiles without a GPU);
# the .so is git-ignored but travels to the GPU box with the snapshot.
_SO = os.path.join(PKG, "fqe_b200", "lib", "libfqe_b200.so")
if not os.path.exists(_SO):
    import importlib.util
    _spec = importlib.util.spec_from_file_location("fqeb_build", os.path.join(PKG, "build.py"))
    _mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
    _mod.build()


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (not failed), so a bare
    ``pytest tests`` on a CPU box is green; the product code itself still raises
    FQEB_ERR_NODEVICE there (tests/test_cabi_symbols.py checks that)."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


# ---------------------------------------------------------------------------------------------
# Two contraction back ends build the same sigma: the INT8-sliced tensor-core kernel (default
# where it applies; digits quantised at 127^-6, measured ~1e-12 relative) and the FP64 DMMA
# kernels (FQEB_OZAKI=0; ~1e-15).  Modules that exercise sigma use the ``contraction`` fixture to
# run on both, each held to its own bound; both are far inside the 1e-10 of BASELINE.json.
# ---------------------------------------------------------------------------------------------
SLICED_TOL = 2.0e-11
FP64_TOL = 1.0e-12


def sigma_tol() -> float:
    """tolerance of the contraction path the environment currently selects"""
    return FP64_TOL if os.environ.get("FQEB_OZAKI") == "0" else SLICED_TOL


@pytest.fixture(params=["sliced", "fp64"])
def contraction(request):
    old = os.environ.get("FQEB_OZAKI")
    if request.param == "fp64":
        os.environ["FQEB_OZAKI"] = "0"
    else:
        os.environ.pop("FQEB_OZAKI", None)
    yield request.param
    if old is None:
        os.environ.pop("FQEB_OZAKI", None)
    else:
        os.environ["FQEB_OZAKI"] = old
