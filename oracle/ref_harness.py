"""ctypes harness over the UNMODIFIED reference C kernels.  TEST INFRASTRUCTURE ONLY.

``oracle/_ref/libfqe_ref.so`` is compiled by ``oracle/Makefile`` straight from
``/root/reference/src/fqe/lib/*.c`` (no source is copied into this repo).  This
module calls those kernels the way the reference's own ctypes / Cython wrappers
do (``lib/fci_graph.py``, ``lib/_fqe_data.pyx``) and restates only the thin
Python glue of ``FqeData._apply_array_spatial12_lm`` (fqe_data.py:685-710) that
sits between them.  It is the strongest oracle available on the GPU box (the
Python reference cannot travel) and the ``cpu_baseline`` of ``bench.py``
(``kind: "reference"``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may
import this module.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_bool, c_int, c_int32, c_uint32, c_uint64, c_void_p

import numpy as np
from scipy.special import comb

_HERE = os.path.dirname(os.path.abspath(__file__))


def _pick_lib_path() -> str:
    """The reference builds with -march=native (setup.py:46); the prebuilt library has
    to travel to another host, so two ISA levels are built and the widest one this
    CPU supports is used."""
    v4 = os.path.join(_HERE, "_ref", "libfqe_ref_v4.so")
    try:
        with open("/proc/cpuinfo") as fh:
            flags = fh.read()
        if os.path.exists(v4) and all(f in flags for f in ("avx512f", "avx512bw", "avx512vl",
                                                           "avx512dq", "avx512cd")):
            return v4
    except OSError:
        pass
    return os.path.join(_HERE, "_ref", "libfqe_ref.so")


_LIB_PATH = _pick_lib_path()

_lib = None
_blas = None
_blas_limit = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{_LIB_PATH} missing: run `make -C oracle` where /root/reference exists")
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.fqe_oracle_blas.restype = c_void_p
        # The reference calls BLAS zaxpy from inside its own OpenMP loops; scipy's
        # OpenBLAS must not spawn threads of its own there (it warns and can
        # oversubscribe for vectors > ~10^4), so pin BLAS to one thread per caller.
        try:
            from threadpoolctl import threadpool_limits
            global _blas_limit
            _blas_limit = threadpool_limits(limits=1, user_api="blas")
        except Exception:  # pragma: no cover
            pass
    return _lib


class _BlasFunctions(ctypes.Structure):
    _fields_ = [("zaxpy", c_void_p), ("zscal", c_void_p)]


def blas_pointer():
    """Address of a `struct blasfunctions` (lib/_blas_helpers.h:33-36).  Like the
    reference's Cython shim (lib/_fqe_data.pyx:83-87) it is filled with scipy's own
    BLAS zaxpy/zscal (scipy.linalg.cython_blas), so the compiled reference runs with
    the same optimised BLAS it ships with; the plain-C routines of ref_blas_shim.c are
    only a fallback when scipy's capsules cannot be read."""
    global _blas
    if _blas is None:
        try:
            from scipy.linalg import cython_blas
            get = ctypes.pythonapi.PyCapsule_GetPointer
            get.restype = c_void_p
            get.argtypes = [ctypes.py_object, ctypes.c_char_p]
            name = ctypes.pythonapi.PyCapsule_GetName
            name.restype = ctypes.c_char_p
            name.argtypes = [ctypes.py_object]
            ptrs = []
            for fn in ("zaxpy", "zscal"):
                cap = cython_blas.__pyx_capi__[fn]
                ptrs.append(get(cap, name(cap)))
            _blas = _BlasFunctions(ptrs[0], ptrs[1])
            _blas.source = "scipy.linalg.cython_blas"
        except Exception:  # pragma: no cover
            _blas = lib().fqe_oracle_blas()
    if isinstance(_blas, _BlasFunctions):
        return ctypes.addressof(_blas)
    return _blas


def _p(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


def _c128(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex128)


class RefGraph:
    """FciGraph products computed by the reference C functions, called as in
    lib/fci_graph.py:31-195 and fci_graph.py:36-63, 301-333."""

    def __init__(self, nalpha: int, nbeta: int, norb: int):
        self.nalpha, self.nbeta, self.norb = nalpha, nbeta, norb
        self.lena = comb(norb, nalpha, exact=True)
        self.lenb = comb(norb, nbeta, exact=True)
        self.za = self._z(nalpha)
        self.zb = self._z(nbeta)
        self.astr = self._strings(nalpha, self.lena, self.za)
        self.bstr = self._strings(nbeta, self.lenb, self.zb)
        self.alpha_map = self._maps(self.astr, self.za)
        self.beta_map = self._maps(self.bstr, self.zb)
        self.dexca = self._dexc(self.alpha_map, self.lena, nalpha)
        self.dexcb = self._dexc(self.beta_map, self.lenb, nbeta)

    def _z(self, nele):
        z = np.zeros((nele, self.norb), dtype=np.int32)
        if z.size:
            f = lib().calculate_Z_matrix
            f.argtypes = [c_void_p, c_int, c_int]
            f.restype = None
            f(_p(z), self.norb, nele)
        return z

    def _strings(self, nele, length, z):
        lex = np.zeros(length, dtype=np.uint64)
        f = lib().lexicographic_bitstring_generator
        f.argtypes = [c_void_p, c_int, c_int]
        f.restype = None
        f(_p(lex), self.norb, nele)
        out = np.zeros(length, dtype=np.uint64)
        g = lib().calculate_string_address
        g.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int]
        g.restype = None
        g(_p(out), _p(lex), length, _p(z), self.norb)
        return out

    def _maps(self, strings, z):
        norb = self.norb
        nmaps = norb * norb
        pairs = np.array([(i, j) for i in range(norb) for j in range(norb)],
                         dtype=np.int32).reshape(-1, 2)
        mapl = np.zeros(nmaps, dtype=np.int32)
        f = lib().build_mapping_strings
        f.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                      c_bool, c_void_p, c_int]
        f.restype = None
        # pass 1: count
        f(None, _p(mapl), _p(pairs), nmaps, _p(strings), strings.shape[0], True,
          _p(z), norb)
        arrays = [np.zeros((int(n), 3), dtype=np.int32) for n in mapl]
        ptrs = (c_void_p * nmaps)(*[a.ctypes.data for a in arrays])
        f(ptrs, _p(mapl), _p(pairs), nmaps, _p(strings), strings.shape[0],
          False, _p(z), norb)
        return {(int(i), int(j)): arrays[k] for k, (i, j) in enumerate(pairs)}

    def _dexc(self, maps, states, nele):
        norb = self.norb
        lk = nele * (norb - nele + 1)
        dexc = np.zeros((states, lk, 3), dtype=np.int32)
        index = np.zeros(states, dtype=np.uint32)
        f = lib().map_deexc
        f.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int]
        f.restype = c_int
        # the reference's map_deexc has an unsynchronised counter that is only
        # safe because targets are unique per (i,j); call as-is.
        for (i, j), m in maps.items():
            if m.shape[0]:
                f(_p(dexc), _p(m), lk, m.shape[0], _p(index), i * norb + j)
        return dexc


_graphs = {}


def graph(nalpha, nbeta, norb) -> RefGraph:
    key = (nalpha, nbeta, norb)
    if key not in _graphs:
        _graphs[key] = RefGraph(nalpha, nbeta, norb)
    return _graphs[key]


def _same_spin(coeff, h1p, h2p, dexc, len1, len2, norb):
    """lib/_fqe_data.pyx:609-680 -> lm_apply_array12_same_spin_opt (fqe_data.c:668)."""
    coeff = _c128(coeff)
    out = np.zeros((len1, len2), dtype=np.complex128)
    f = lib().lm_apply_array12_same_spin_opt
    f.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                  c_void_p, c_int, c_bool, c_void_p]
    f.restype = None
    f(_p(coeff), _p(out), _p(dexc), len1, len2, dexc.shape[1], _p(h1p),
      _p(h2p), norb, True, blas_pointer())
    return out


def sigma_restricted(g: RefGraph, coeff, h1, h2, parts=None) -> np.ndarray:
    """FqeData._apply_array_spatial12_lm (fqe_data.py:685-710): production C path."""
    norb = g.norb
    coeff = _c128(coeff)
    h2p = _c128(-np.moveaxis(np.asarray(h2, dtype=np.complex128), 1, 2))
    h1p = _c128(np.asarray(h1, dtype=np.complex128) -
                np.einsum("ikkj->ij", h2p))
    out = _same_spin(coeff, h1p, h2p, g.dexca, g.lena, g.lenb, norb)
    out += _same_spin(coeff.T, h1p, h2p, g.dexcb, g.lenb, g.lena, norb).T
    h2d = _c128(h2p + np.einsum("ijkl->klij", h2p))
    f = lib().lm_apply_array12_diff_spin_opt
    f.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                  c_int, c_void_p, c_int, c_void_p]
    f.restype = None
    out = _c128(out)
    f(_p(coeff), _p(out), _p(g.dexca), _p(g.dexcb), g.lena, g.lenb,
      g.dexca.shape[1], g.dexcb.shape[1], _p(h2d), norb,
      blas_pointer())
    return out


def _make(fn_name, g: RefGraph, maps_a, maps_b, src, dst):
    """zdvec_make / zcoeff_make, called as lib/_fqe_data.pyx:408-505 does."""
    f = getattr(lib(), fn_name)
    f.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                  c_bool, c_void_p]
    f.restype = c_int
    for maps, is_alpha in ((maps_a, True), (maps_b, False)):
        n = len(maps)
        els = np.array([m.shape[0] for m in maps], dtype=np.int32)
        keep = [np.ascontiguousarray(m, dtype=np.int32) for m in maps]
        ptrs = (c_void_p * n)(*[m.ctypes.data for m in keep])
        f(ptrs, _p(els), n, _p(src), _p(dst), g.lena, g.lenb, is_alpha,
          blas_pointer())


def dvec_spatial(g: RefGraph, coeff) -> np.ndarray:
    """calculate_dvec_spatial through zdvec_make (fqe_data.py:2209-2224)."""
    n = g.norb
    coeff = _c128(coeff)
    dvec = np.zeros((n, n, g.lena, g.lenb), dtype=np.complex128)
    ma = [g.alpha_map[(i, j)] for i in range(n) for j in range(n)]
    mb = [g.beta_map[(i, j)] for i in range(n) for j in range(n)]
    # zdvec_make(map, els, n, coeff, dvec, ...)
    _make("zdvec_make", g, ma, mb, coeff, dvec)
    return dvec


def coeff_from_dvec(g: RefGraph, dvec) -> np.ndarray:
    """_calculate_coeff_spatial_with_dvec through zcoeff_make (fqe_data.py:2317-2327)."""
    n = g.norb
    dvec = _c128(dvec)
    out = np.zeros((g.lena, g.lenb), dtype=np.complex128)
    ma = [g.alpha_map[(j, i)] for i in range(n) for j in range(n)]
    mb = [g.beta_map[(j, i)] for i in range(n) for j in range(n)]
    # zcoeff_make(map, els, n, coeff(out), dvec(in), ...)
    _make("zcoeff_make", g, ma, mb, out, dvec)
    return out


def _dc(fn_name, g: RefGraph, coeff, diag, array):
    out = _c128(coeff).copy()
    diag = _c128(diag)
    array = _c128(array)
    f = getattr(lib(), fn_name)
    f.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                  c_int, c_int, c_int, c_int]
    f.restype = None
    f(_p(g.astr), _p(g.bstr), _p(diag), _p(array), _p(out), g.lena, g.lenb,
      g.nalpha, g.nbeta, g.norb)
    return out


def dc_apply(g, coeff, diag, array):
    """zdiagonal_coulomb_apply (lib/fqe_data.c:455-524)."""
    return _dc("zdiagonal_coulomb_apply", g, coeff, diag, array)


def dc_evolve(g, coeff, diag, array):
    """zdiagonal_coulomb (lib/fqe_data.c:526-602)."""
    return _dc("zdiagonal_coulomb", g, coeff, diag, array)


# ---------------------------------------------------------------------------
# bounded, exact-work sampling of the reference sigma for CPU timing
# ---------------------------------------------------------------------------
def time_sigma_sample(g: RefGraph, coeff, h1, h2, m: int):
    """Time the three reference kernels of one sigma build on an exact 1/f slice of
    their work and return (seconds_per_kernel dict, f = lena / m).

    Each kernel's cost is linear in one index that can be truncated without
    touching the reference code:
      * same-spin alpha  (lm_apply_array12_same_spin_opt on C): linear in the number
        of beta columns (the zaxpy length, lib/fqe_data.c:711-715) -> first m columns;
      * same-spin beta   (same kernel on C^T): linear in alpha rows -> first m rows;
      * opposite-spin    (lm_apply_array12_diff_spin_opt): linear in the alpha strings
        scanned per orbital pair (lib/fqe_data.c:911-939) -> dexc of the first m alpha
        strings; requires m <= C(norb-1, nalpha-1) so the scratch sizing by `nest`
        (fqe_data.c:903-908) stays an upper bound.
    The outputs are partial and NOT a sigma vector; this is a stopwatch only.
    """
    import time
    norb = g.norb
    coeff = _c128(coeff)
    m = int(min(m, g.lena, g.lenb, comb(norb - 1, g.nalpha - 1, exact=True) if g.nalpha else 1))
    m = max(m, 1)
    h2p = _c128(-np.moveaxis(np.asarray(h2, dtype=np.complex128), 1, 2))
    h1p = _c128(np.asarray(h1, dtype=np.complex128) - np.einsum("ikkj->ij", h2p))
    h2d = _c128(h2p + np.einsum("ijkl->klij", h2p))
    times = {}
    ca = _c128(coeff[:, :m])
    t0 = time.perf_counter()
    _same_spin(ca, h1p, h2p, g.dexca, g.lena, m, norb)
    times["same_spin_alpha"] = time.perf_counter() - t0
    cb = _c128(coeff[:m, :].T)
    t0 = time.perf_counter()
    _same_spin(cb, h1p, h2p, g.dexcb, g.lenb, m, norb)
    times["same_spin_beta"] = time.perf_counter() - t0
    f = lib().lm_apply_array12_diff_spin_opt
    f.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                  c_int, c_void_p, c_int, c_void_p]
    f.restype = None
    out = np.zeros((m, g.lenb), dtype=np.complex128)
    adexc = np.ascontiguousarray(g.dexca[:m])
    t0 = time.perf_counter()
    f(_p(coeff), _p(out), _p(adexc), _p(g.dexcb), m, g.lenb, adexc.shape[1],
      g.dexcb.shape[1], _p(h2d), norb, blas_pointer())
    times["diff_spin"] = time.perf_counter() - t0
    return times, g.lena / m


def estimate_sigma_seconds(g: RefGraph, coeff, h1, h2, budget_s: float = 15.0):
    """Wall time of ONE full reference sigma build, from exact-work samples.

    Every kernel's time is affine in the truncation length m (a fixed per-call
    overhead plus work proportional to m, see time_sigma_sample), so two sample
    sizes give t(m) = a + b*m per kernel and the full cost is sum_k a_k + b_k*lena.
    Sample sizes are grown until the larger one costs about ``budget_s`` seconds.
    Returns (seconds, description)."""
    mmax = int(min(g.lena, g.lenb, comb(g.norb - 1, g.nalpha - 1, exact=True) if g.nalpha else 1))
    if mmax >= g.lena or g.lena <= 64:
        import time
        t0 = time.perf_counter()
        sigma_restricted(g, coeff, h1, h2)
        return time.perf_counter() - t0, "one full sigma build"
    m_lo = max(2, min(16, mmax // 4))
    t_lo, _ = time_sigma_sample(g, coeff, h1, h2, m_lo)
    m_hi = min(mmax, 4 * m_lo)
    t_hi, _ = time_sigma_sample(g, coeff, h1, h2, m_hi)

    def fit(ta, ma, tb, mb):
        total = 0.0
        for k in ta:
            b = max((tb[k] - ta[k]) / (mb - ma), 0.0)
            a = max(tb[k] - b * mb, 0.0)
            total += a + b * g.lena
        return total

    full = fit(t_lo, m_lo, t_hi, m_hi)
    m_big = int(g.lena * budget_s / max(full, 1e-9))
    m_big = max(min(m_big, mmax), min(mmax, 2 * m_hi))
    if m_big > m_hi:
        t_big, _ = time_sigma_sample(g, coeff, h1, h2, m_big)
        full = fit(t_hi, m_hi, t_big, m_big)
        m_lo, m_hi = m_hi, m_big
    desc = (f"3 reference kernels timed on the first {m_lo} and {m_hi} of {g.lena} "
            f"alpha rows / beta columns (exact 1/f work slices), affine fit per kernel "
            f"extrapolated to the full sigma")
    return full, desc
