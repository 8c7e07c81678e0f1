"""ctypes binding of libfqe_b200.so (the C ABI declared in include/fqe_b200.h).

Mirrors how the reference loads its native library
(/root/reference/src/fqe/lib/__init__.py:4-14: ``ctypes.cdll.LoadLibrary`` of a
``.so`` that sits next to the package).  There is NO fallback: if the shared
object is missing, or no CUDA device is visible when a compute entry point is
called, an exception is raised.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint64,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
# FQEB_B200_LIB points at an alternative build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("FQEB_B200_LIB") or os.path.join(_HERE, "libfqe_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_NODEVICE, ERR_CONVERGE = range(6)
OP_REAL, OP_IMAG, OP_COMPLEX = 0, 1, 2
OP_FLAG_FULL_PAIR_SPACE = 1


class FqeB200Error(RuntimeError):
    """A libfqe_b200 entry point returned a non-zero status."""

    def __init__(self, code, message):
        super().__init__(f"libfqe_b200 error {code}: {message}")
        self.code = code


class PendingScatter(ctypes.Structure):
    """``fqeb_pending_scatter``: the deferred scatter of the last chunk of a sigma build"""
    _fields_ = [("d_evec", c_void_p), ("lde", c_int64), ("pitch", c_int64), ("row0", c_int64),
                ("nrows", c_int64), ("d_rowmap", c_void_p), ("zr", c_double), ("zi", c_double)]


# name -> (restype, argtypes); every symbol include/fqe_b200.h declares
SIGNATURES = {
    "fqeb_last_error": (c_char_p, []),
    "fqeb_version": (c_int, []),
    "fqeb_device_count": (c_int, []),
    "fqeb_launch_count": (c_uint64, []),
    "fqeb_sigma_last_path": (c_int, []),
    "fqeb_host_release": (c_int, []),
    "fqeb_i8_tensor_peak": (c_int, [POINTER(c_double)]),
    "fqeb_ozaki_profile": (c_int, [POINTER(c_uint64)]),
    "fqeb_set_device": (c_int, [c_int]),
    "fqeb_graph_create": (c_int, [c_int, c_int, c_int, POINTER(c_void_p)]),
    "fqeb_graph_destroy": (c_int, [c_void_p]),
    "fqeb_graph_dims": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                POINTER(c_int64), POINTER(c_int64)]),
    "fqeb_graph_get_Z": (c_int, [c_void_p, c_int, c_void_p]),
    "fqeb_graph_get_strings": (c_int, [c_void_p, c_int, c_void_p]),
    "fqeb_graph_get_map": (c_int, [c_void_p, c_int, c_void_p]),
    "fqeb_graph_device_tables": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p),
                                         POINTER(c_void_p)]),
    "fqeb_op_create": (c_int, [c_int, c_void_p, c_void_p, POINTER(c_void_p)]),
    "fqeb_op_create_ex": (c_int, [c_int, c_void_p, c_void_p, c_int, POINTER(c_void_p)]),
    "fqeb_op_destroy": (c_int, [c_void_p]),
    "fqeb_op_destroy_async": (c_int, [c_void_p, c_void_p]),
    "fqeb_op_kind": (c_int, [c_void_p, POINTER(c_int)]),
    "fqeb_op_pair_space": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    "fqeb_make_dvec": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int,
                               c_int, c_void_p]),
    "fqeb_make_coeff": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_double, c_double,
                                c_void_p, c_void_p]),
    "fqeb_contract": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int,
                              c_int, c_void_p]),
    "fqeb_gemm_col_align": (c_int, []),
    "fqeb_gram_accumulate": (c_int, [c_int, c_int, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                     c_void_p, c_void_p, c_void_p]),
    "fqeb_contract_dvec_rows": (c_int, [c_void_p, c_int]),
    "fqeb_sigma_workspace_bytes": (c_size_t, [c_void_p, c_void_p, c_int64, c_int, c_int]),
    "fqeb_sigma_rows_for_workspace": (c_int64, [c_void_p, c_void_p, c_size_t, c_int, c_int]),
    "fqeb_sigma_restricted": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_int64, c_int64, c_int, c_int, c_void_p]),
    "fqeb_sigma_restricted_deferred": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_size_t, c_int64, c_int64, c_int, c_int,
                                               POINTER(PendingScatter), c_void_p]),
    "fqeb_scatter_rows": (c_int, [c_void_p, POINTER(PendingScatter), c_int64, c_int64, c_void_p,
                                  c_void_p]),
    "fqeb_sigma_restricted_host": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                           c_void_p]),
    "fqeb_profile_enable": (c_int, [c_int]),
    "fqeb_profile_collect": (c_int, [POINTER(c_double), POINTER(c_int64)]),
    "fqeb_dc_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fqeb_dc_evolve": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fqeb_apply_columns": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "fqeb_apply_diagonal": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fqeb_evolve_diagonal": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fqeb_nbody_accumulate": (c_int, [c_void_p, c_double, c_double, c_void_p, c_void_p, c_int,
                                      c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "fqeb_sparse_scale": (c_int, [c_void_p, c_uint64, c_uint64, c_uint64, c_uint64, c_double,
                                  c_double, c_void_p, c_void_p]),
    "fqeb_taylor": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                            c_void_p, c_double, c_int, POINTER(c_int), c_void_p]),
    "fqeb_reduce_scratch_bytes": (c_size_t, []),
    "fqeb_zaxpy": (c_int, [c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "fqeb_zscal": (c_int, [c_int64, c_double, c_double, c_void_p, c_void_p]),
    "fqeb_zaxpby": (c_int, [c_int64, c_double, c_double, c_void_p, c_double, c_double, c_void_p,
                            c_void_p]),
    "fqeb_znorm2": (c_int, [c_int64, c_void_p, c_void_p, POINTER(c_double), c_void_p]),
    "fqeb_zdotc": (c_int, [c_int64, c_void_p, c_void_p, c_void_p, POINTER(c_double), c_void_p]),
    "fqeb_axpy_norm2": (c_int, [c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                POINTER(c_double), c_void_p]),
}

_lib = None


def load():
    """Load libfqe_b200.so (once) and attach the signatures.  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FqeB200Error(
                -1, f"{LIB_PATH} not found: build it with `python openfermion-fqe_b200/build.py` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code):
    if code != OK:
        raise FqeB200Error(code, load().fqeb_last_error().decode("utf-8", "replace"))


def call(name, *args):
    """Call a status-returning entry point and raise on failure."""
    check(getattr(load(), name)(*args))
