"""FqeData: one (n, sz) sector of a wavefunction, resident in HBM.

Host-side mirror of the reference class (/root/reference/src/fqe/fqe_data.py:60-
3076) for the Hamiltonian-application path.  ``coeff`` is a ``torch.complex128``
CUDA tensor of shape ``[lena, lenb]`` (row = alpha string, column = beta string,
C-contiguous, exactly the reference's numpy layout, fqe_data.py:106) and every
method below forwards to a hand-written sm_100a kernel through the C ABI of
libfqe_b200.so.  There is no CPU code path.

Method names, argument meaning and error behaviour follow the reference so the
parity tests read like its own tests (tests/fqe_data_test.py).
"""
import copy
import ctypes
import math
from typing import Optional, Tuple

import numpy
import torch

from fqe_b200 import lib as _lib
from fqe_b200 import settings
from fqe_b200.fci_graph import FciGraph, get_graph

_C128 = numpy.complex128


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.FqeB200Error(_lib.ERR_NODEVICE,
                                "no CUDA device visible; fqe_b200 has no CPU fallback")
    dev = torch.cuda.current_device()
    _lib.call("fqeb_set_device", dev)
    return torch.device("cuda", dev)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _host_c128(arr) -> numpy.ndarray:
    return numpy.ascontiguousarray(numpy.asarray(arr), dtype=_C128)


def validate_config(nalpha: int, nbeta: int, norb: int) -> None:
    """Same checks as util.validate_config in the reference (util.py)."""
    if nalpha < 0:
        raise ValueError("Cannot have negative number of alpha electrons")
    if nbeta < 0:
        raise ValueError("Cannot have negative number of beta electrons")
    if norb < 0:
        raise ValueError("Cannot have negative number of orbitals")
    if norb < nalpha or norb < nbeta:
        raise ValueError("Insufficient number of orbitals")


# ---------------------------------------------------------------------------
# device scratch shared by all sectors on a device
# ---------------------------------------------------------------------------
_SCRATCH = {}
_WORKSPACE = {}
_WS_CAPPED = set()


def _reduce_scratch(dev: torch.device) -> torch.Tensor:
    if dev not in _SCRATCH:
        nbytes = int(_lib.load().fqeb_reduce_scratch_bytes())
        _SCRATCH[dev] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    return _SCRATCH[dev]


def _workspace(dev: torch.device, wanted: int, minimum: int, reserve: int = 0) -> torch.Tensor:
    """A byte workspace of at most ``wanted`` and at least ``minimum`` bytes, bounded by
    settings.workspace_fraction of the free memory and leaving ``reserve`` bytes (one more
    coefficient vector) unclaimed.  Cached (and grown on demand) so that repeated sigma
    builds - the Taylor loop - never reallocate."""
    cur = _WORKSPACE.get(dev)
    if cur is not None and (cur.numel() >= wanted or
                            (dev in _WS_CAPPED and cur.numel() >= minimum)):
        return cur
    have = cur.numel() if cur is not None else 0
    _WORKSPACE.pop(dev, None)
    _WS_CAPPED.discard(dev)
    del cur
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info(dev)
    budget = min(int(free * settings.workspace_fraction), free - reserve - (1 << 30))
    if settings.max_workspace_bytes is not None:
        budget = min(budget, int(settings.max_workspace_bytes))
    size = min(wanted, max(budget, minimum, have))
    if size < minimum:
        raise _lib.FqeB200Error(_lib.ERR_NOMEM,
                                f"sigma workspace needs at least {minimum} bytes, budget is {size}")
    ws = torch.empty(size, dtype=torch.uint8, device=dev)
    _WORKSPACE[dev] = ws
    if size < wanted:
        _WS_CAPPED.add(dev)   # as large as the budget allows: keep it
    return ws


def release_workspace() -> None:
    """Free the cached sigma workspace (all devices)."""
    _WORKSPACE.clear()
    _WS_CAPPED.clear()
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------
# prepared dense operator
# ---------------------------------------------------------------------------
def fold_restricted(h1e: numpy.ndarray, h2e: numpy.ndarray) -> Tuple[numpy.ndarray, numpy.ndarray]:
    """Tensor preparation of FqeData._apply_array_spatial12_lm (reference fqe_data.py:691-693):
    h2' = -moveaxis(h2, 1, 2), h1' = h1 - sum_k h2'[i,k,k,j]; complex128, C-contiguous - the
    arguments ``fqeb_op_create`` and ``fqeb_sigma_restricted_host`` take."""
    h2p = numpy.ascontiguousarray(-numpy.moveaxis(numpy.asarray(h2e).astype(_C128), 1, 2))
    h1p = numpy.ascontiguousarray(numpy.asarray(h1e).astype(_C128) -
                                  numpy.einsum("ikkj->ij", h2p))
    return h1p, h2p


class DenseOperator:
    """(h1, h2) folded and uploaded for repeated sigma builds.

    Performs the tensor preparation of FqeData._apply_array_spatial12_lm
    (fqe_data.py:691-693): h2' = -moveaxis(h2, 1, 2), h1' = h1 - sum_k h2'[i,k,k,j]
    on the host (O(norb^4)), then hands both to ``fqeb_op_create``.
    """

    def __init__(self, norb: int, h1e: numpy.ndarray, h2e: Optional[numpy.ndarray] = None):
        _require_cuda()
        h1e = numpy.asarray(h1e)
        if h1e.shape != (norb, norb):
            raise ValueError(f"h1e has shape {h1e.shape}, expected {(norb, norb)}")
        h1p = h1e.astype(_C128, copy=True)
        h2p_ptr = None
        self._h2p = None
        if h2e is not None:
            h2e = numpy.asarray(h2e)
            if h2e.shape != (norb,) * 4:
                raise ValueError(f"h2e has shape {h2e.shape}, expected {(norb,) * 4}")
            h2p = numpy.ascontiguousarray(-numpy.moveaxis(h2e.astype(_C128), 1, 2))
            h1p -= numpy.einsum("ikkj->ij", h2p)
            tol = float(settings.symmetry_tolerance)
            if tol > 0.0:
                avg = 0.25 * (h2p + h2p.transpose(1, 0, 2, 3) + h2p.transpose(0, 1, 3, 2) +
                              h2p.transpose(1, 0, 3, 2))
                scale = float(numpy.abs(h2p).max())
                if 0.0 < float(numpy.abs(avg - h2p).max()) <= tol * scale:
                    h2p = numpy.ascontiguousarray(avg)
            self._h2p = h2p
            h2p_ptr = h2p.ctypes.data
        h1p = numpy.ascontiguousarray(h1p)
        self.norb = norb
        self.has_h2 = h2e is not None
        handle = ctypes.c_void_p()
        _lib.call("fqeb_op_create", norb, h1p.ctypes.data, h2p_ptr, ctypes.byref(handle))
        self._handle = handle
        kind = ctypes.c_int()
        _lib.call("fqeb_op_kind", handle, ctypes.byref(kind))
        self.kind = int(kind.value)
        npairs, sym = ctypes.c_int(), ctypes.c_int()
        _lib.call("fqeb_op_pair_space", handle, ctypes.byref(npairs), ctypes.byref(sym))
        self.npair = int(npairs.value)      # pair space of the contraction
        self.symmetric = bool(sym.value)    # i>=j compressed (real-orbital integrals)

    @classmethod
    def from_folded(cls, norb: int, h1p: numpy.ndarray, h2p: Optional[numpy.ndarray],
                    full_pair_space: bool = False):
        """Operator from ALREADY folded tensors: E[ij] = sum_kl h2p[i,j,k,l] D[k,l] and
        sigma += sum_ij h1p[i,j] D[i,j] (no moveaxis / trace folding applied).
        ``full_pair_space`` keeps the norb^2 pair space even for pair-symmetric tensors."""
        self = cls.__new__(cls)
        _require_cuda()
        h1p = numpy.ascontiguousarray(numpy.asarray(h1p), dtype=_C128)
        self._h2p = None
        ptr = None
        if h2p is not None:
            self._h2p = numpy.ascontiguousarray(numpy.asarray(h2p), dtype=_C128)
            ptr = self._h2p.ctypes.data
        self.norb = norb
        self.has_h2 = h2p is not None
        handle = ctypes.c_void_p()
        _lib.call("fqeb_op_create_ex", norb, h1p.ctypes.data, ptr,
                  _lib.OP_FLAG_FULL_PAIR_SPACE if full_pair_space else 0, ctypes.byref(handle))
        self._handle = handle
        kind, npairs, sym = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.call("fqeb_op_kind", handle, ctypes.byref(kind))
        _lib.call("fqeb_op_pair_space", handle, ctypes.byref(npairs), ctypes.byref(sym))
        self.kind, self.npair, self.symmetric = int(kind.value), int(npairs.value), bool(sym.value)
        return self

    @property
    def handle(self) -> ctypes.c_void_p:
        """Native handle.  Every access notes the current CUDA stream: the device buffers are
        released in that stream's order when the operator is dropped, i.e. after the kernels
        that were handed this handle, without synchronising the device."""
        self._streams = getattr(self, "_streams", set())
        if torch.cuda.is_available():
            self._streams.add(_stream())
        return self._handle

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                streams = getattr(self, "_streams", set())
                if len(streams) == 1:
                    _lib.load().fqeb_op_destroy_async(h, next(iter(streams)))
                else:   # never used, or used on several streams: wait for the device
                    _lib.load().fqeb_op_destroy(h)
            except Exception:
                pass
            self._handle = None


def op_factor(op: DenseOperator) -> complex:
    """Scalar the contraction factors out of the operator: i for purely imaginary
    tensors (their imaginary part is contracted as a real matrix), else 1."""
    return 1.0j if op.kind == _lib.OP_IMAG else 1.0 + 0.0j


class FqeData:
    """Coefficients of one sector plus the kernels that act on them."""

    def __init__(self,
                 nalpha: int,
                 nbeta: int,
                 norb: int,
                 fcigraph: Optional[FciGraph] = None,
                 dtype=numpy.complex128) -> None:
        validate_config(nalpha, nbeta, norb)
        if fcigraph is not None and (nalpha != fcigraph.nalpha() or nbeta != fcigraph.nbeta() or
                                     norb != fcigraph.norb()):
            raise ValueError("FciGraph does not match other parameters")
        if numpy.dtype(dtype) != numpy.dtype(numpy.complex128):
            raise TypeError("fqe_b200 stores coefficients as complex128 only")
        dev = _require_cuda()
        self._core = fcigraph if fcigraph is not None else get_graph(nalpha, nbeta, norb)
        self._dtype = numpy.complex128
        self._nele = nalpha + nbeta
        self._m_s = nalpha - nbeta
        self.coeff = torch.zeros((self.lena(), self.lenb()), dtype=torch.complex128, device=dev)

    # ---- bookkeeping (fqe_data.py:2536-2618) ---------------------------------------
    def get_fcigraph(self) -> FciGraph:
        return self._core

    def n_electrons(self) -> int:
        return self._nele

    def nalpha(self) -> int:
        return self._core.nalpha()

    def nbeta(self) -> int:
        return self._core.nbeta()

    def norb(self) -> int:
        return self._core.norb()

    def lena(self) -> int:
        return self._core.lena()

    def lenb(self) -> int:
        return self._core.lenb()

    def ndim(self) -> int:
        return 2

    def alpha_map(self, iorb: int, jorb: int):
        return self._core.alpha_map(iorb, jorb)

    def beta_map(self, iorb: int, jorb: int):
        return self._core.beta_map(iorb, jorb)

    def __hash__(self):
        return hash((self._nele, self._m_s))

    def __getitem__(self, key: Tuple[int, int]) -> complex:
        return complex(self.coeff[self._core.index_alpha(key[0]),
                                  self._core.index_beta(key[1])].item())

    def __setitem__(self, key: Tuple[int, int], value: complex) -> None:
        self.coeff[self._core.index_alpha(key[0]), self._core.index_beta(key[1])] = value

    def __deepcopy__(self, memodict={}) -> 'FqeData':
        new = self.empty_copy(zero=False)
        new.coeff.copy_(self.coeff)
        return new

    def empty_copy(self, zero: bool = True) -> 'FqeData':
        new = FqeData.__new__(FqeData)
        new._core = self._core
        new._dtype = self._dtype
        new._nele = self._nele
        new._m_s = self._m_s
        new.coeff = torch.zeros_like(self.coeff) if zero else torch.empty_like(self.coeff)
        return new

    def to_numpy(self) -> numpy.ndarray:
        """Host copy of the coefficients (numpy complex128 [lena, lenb])."""
        return self.coeff.detach().cpu().numpy()

    # ---- initialisation (fqe_data.py:2761-2810) ----------------------------------
    def set_wfn(self, strategy: Optional[str] = None, raw_data=None) -> None:
        strategy_args = ['ones', 'zero', 'random', 'from_data', 'hartree-fock']
        no_data = raw_data is None or (hasattr(raw_data, "shape") and tuple(raw_data.shape) == (0,))
        if strategy is None and no_data:
            raise ValueError('No strategy and no data passed. Cannot initialize')
        if strategy == 'from_data' and no_data:
            raise ValueError('No data passed to initialize from')
        if not no_data and strategy not in ['from_data', None]:
            raise ValueError('Inconsistent strategy for set_vec passed with data')
        if strategy is None:
            strategy = 'from_data'
        if strategy not in strategy_args:
            raise ValueError('Unknown Argument passed to set_vec')
        if strategy == 'from_data':
            shape = tuple(raw_data.shape)
            if len(shape) != 2 or shape[0] != self.lena() or shape[1] != self.lenb():
                raise ValueError('Dim of data passed {} is not compatible with {},{}'.format(
                    shape, self.lena(), self.lenb()))
            if isinstance(raw_data, torch.Tensor):
                self.coeff.copy_(raw_data.to(torch.complex128))
            else:
                host = torch.from_numpy(_host_c128(raw_data))
                self.coeff.copy_(host)
        elif strategy == 'ones':
            self.coeff.fill_(1.0 + 0.0j)
        elif strategy == 'zero':
            self.coeff.zero_()
        elif strategy == 'random':
            # util.rand_wfn: unnormalised complex standard normal (util.py:383-397)
            shp = (self.lena(), self.lenb())
            host = numpy.random.randn(*shp) + 1.0j * numpy.random.randn(*shp)
            self.coeff.copy_(torch.from_numpy(host))
        elif strategy == 'hartree-fock':
            self.coeff.zero_()
            self.coeff[0, 0] = 1.0

    def fill(self, value: complex) -> None:
        self.coeff.fill_(value)

    def conj(self) -> None:
        self.coeff = torch.conj_physical(self.coeff)

    # ---- BLAS-1 (fqe_data.py:2620-2632, 2701-2707, 2745-2751) -----------------------
    def _n(self) -> int:
        return self.coeff.numel()

    def _check_coeff(self, t: torch.Tensor) -> torch.Tensor:
        if not (t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()):
            raise TypeError("coefficients must be a contiguous CUDA complex128 tensor")
        return t

    def ax_plus_y(self, sval: complex, other: 'FqeData') -> 'FqeData':
        """self.coeff += sval * other.coeff"""
        assert hash(self) == hash(other)
        _require_cuda()
        sval = complex(sval)
        _lib.call("fqeb_zaxpy", self._n(), sval.real, sval.imag,
                  self._check_coeff(other.coeff).data_ptr(),
                  self._check_coeff(self.coeff).data_ptr(), _stream())
        return self

    def scale(self, sval: complex) -> None:
        _require_cuda()
        sval = complex(sval)
        _lib.call("fqeb_zscal", self._n(), sval.real, sval.imag,
                  self._check_coeff(self.coeff).data_ptr(), _stream())

    def norm(self) -> float:
        dev = _require_cuda()
        out = ctypes.c_double()
        _lib.call("fqeb_znorm2", self._n(), self._check_coeff(self.coeff).data_ptr(),
                  _reduce_scratch(dev).data_ptr(), ctypes.byref(out), _stream())
        return math.sqrt(out.value)

    def vdot(self, other: 'FqeData') -> complex:
        """sum conj(self) * other (util.vdot, util.py:506-530)."""
        dev = _require_cuda()
        out = (ctypes.c_double * 2)()
        _lib.call("fqeb_zdotc", self._n(), self._check_coeff(self.coeff).data_ptr(),
                  self._check_coeff(other.coeff).data_ptr(), _reduce_scratch(dev).data_ptr(),
                  out, _stream())
        return complex(out[0], out[1])

    def axpy_norm(self, sval: complex, work: 'FqeData') -> float:
        """self += sval*work and return ||work|| in one pass over memory (the body of
        the Taylor loop, wavefunction.py:563-566)."""
        dev = _require_cuda()
        sval = complex(sval)
        out = ctypes.c_double()
        _lib.call("fqeb_axpy_norm2", self._n(), sval.real, sval.imag,
                  self._check_coeff(work.coeff).data_ptr(),
                  self._check_coeff(self.coeff).data_ptr(), _reduce_scratch(dev).data_ptr(),
                  ctypes.byref(out), _stream())
        return math.sqrt(out.value)

    # ---- dense operator application (fqe_data.py:404-475) -------------------------
    def apply(self, array: Tuple[numpy.ndarray, ...]) -> 'FqeData':
        out = copy.deepcopy(self)
        out.apply_inplace(array)
        return out

    def apply_inplace(self, array: Tuple[numpy.ndarray, ...]) -> None:
        len_arr = len(array)
        if len_arr < 1 or len_arr > 4:
            raise ValueError("Number of operators in tuple must be between 1 and 4.")
        first = next((x for x in array if isinstance(x, numpy.ndarray)), None)
        if first is None:
            return
        spatial = first.shape[0] == self.norb()
        if not spatial and first.shape[0] != 2 * self.norb():
            raise ValueError("Inconsistent number of spin-orbitals in operators and wavefunction.")
        if not spatial:
            raise NotImplementedError(
                "fqe_b200 accelerates spatial-orbital (restricted) operators only")
        if len_arr == 1:
            self.coeff = self._apply_array_spatial1(array[0])
        elif len_arr == 2:
            self.coeff = self._apply_array_spatial12(array[0], array[1])
        elif len_arr == 3:
            self.coeff = self._apply_array_spatial123(array[0], array[1], array[2])
        else:
            raise NotImplementedError("4-body dense operators are outside the B200 hot path")

    def apply_operator(self, op: DenseOperator, row_range=None, pair_range=None,
                       out: Optional[torch.Tensor] = None, defer_last_scatter: bool = False):
        """sigma for a prepared operator.  ``row_range`` / ``pair_range`` restrict the
        work to one rank's shard (partial sigma); ``out`` is an optional caller-owned
        contiguous complex128 [lena, lenb] CUDA tensor that receives the result.
        With ``defer_last_scatter`` the scatter of the last chunk is left to the caller:
        returns ``(sigma, pending)`` and ``finish_scatter(pending, x0, x1, sigma)`` completes
        target rows [x0, x1) (multi-GPU: the all-reduce of finished rows overlaps the rest)."""
        dev = _require_cuda()
        if op.norb != self.norb():
            raise ValueError("operator / wavefunction orbital mismatch")
        r0, r1 = row_range if row_range is not None else (0, self.lena())
        p0, p1 = pair_range if pair_range is not None else (0, op.npair)
        if out is None:
            sigma = torch.empty_like(self.coeff)
        else:
            if tuple(out.shape) != tuple(self.coeff.shape):
                raise ValueError("out has the wrong shape")
            sigma = self._check_coeff(out)
        ws_ptr, ws_bytes = None, 0
        if op.has_h2 and r1 > r0 and p1 > p0:
            lib = _lib.load()
            wanted = int(lib.fqeb_sigma_workspace_bytes(self._core.handle, op.handle, r1 - r0,
                                                        p0, p1))
            minimum = int(lib.fqeb_sigma_workspace_bytes(self._core.handle, op.handle, 1, p0, p1))
            ws = _workspace(dev, wanted, minimum, reserve=16 * self._n())
            ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
        if defer_last_scatter:
            pending = _lib.PendingScatter()
            _lib.call("fqeb_sigma_restricted_deferred", self._core.handle, op.handle,
                      self._check_coeff(self.coeff).data_ptr(), sigma.data_ptr(), ws_ptr,
                      ws_bytes, r0, r1, p0, p1, ctypes.byref(pending), _stream())
            return sigma, pending
        _lib.call("fqeb_sigma_restricted", self._core.handle, op.handle,
                  self._check_coeff(self.coeff).data_ptr(), sigma.data_ptr(), ws_ptr, ws_bytes,
                  r0, r1, p0, p1, _stream())
        return sigma

    def finish_scatter(self, pending, x0: int, x1: int, sigma: torch.Tensor) -> None:
        """complete target rows [x0, x1) of a sigma built with ``defer_last_scatter``"""
        _lib.call("fqeb_scatter_rows", self._core.handle, ctypes.byref(pending), int(x0), int(x1),
                  self._check_coeff(sigma).data_ptr(), _stream())

    def taylor_inplace(self, op: DenseOperator, accuracy: float = 1.0e-15,
                       expansion: int = 30) -> int:
        """coeff <- sum_k op^k coeff / k! with the whole recurrence in one native call
        (``fqeb_taylor``); ``op`` is prepared from the tensors of -i*t*H.  Returns the number of
        terms; raises RuntimeError when ``expansion`` is reached (wavefunction.py:548-567)."""
        dev = _require_cuda()
        if op.norb != self.norb():
            raise ValueError("operator / wavefunction orbital mismatch")
        lib = _lib.load()
        ws_ptr, ws_bytes = None, 0
        if op.has_h2:
            wanted = int(lib.fqeb_sigma_workspace_bytes(self._core.handle, op.handle, self.lena(),
                                                        0, op.npair))
            minimum = int(lib.fqeb_sigma_workspace_bytes(self._core.handle, op.handle, 1, 0,
                                                         op.npair))
            ws = _workspace(dev, wanted, minimum, reserve=32 * self._n())
            ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
        work, nxt = torch.empty_like(self.coeff), torch.empty_like(self.coeff)
        nterms = ctypes.c_int()
        code = lib.fqeb_taylor(self._core.handle, op.handle,
                               self._check_coeff(self.coeff).data_ptr(), work.data_ptr(),
                               nxt.data_ptr(), ws_ptr, ws_bytes, _reduce_scratch(dev).data_ptr(),
                               float(accuracy), int(expansion), ctypes.byref(nterms), _stream())
        if code == _lib.ERR_CONVERGE:
            raise RuntimeError("maximum taylor expansion limit reached")
        _lib.check(code)
        return int(nterms.value)

    def _apply_array_spatial1(self, h1e: numpy.ndarray) -> torch.Tensor:
        assert h1e.shape == (self.norb(), self.norb())
        return self.apply_operator(DenseOperator(self.norb(), h1e, None))

    def _apply_array_spatial12(self, h1e: numpy.ndarray, h2e: numpy.ndarray) -> torch.Tensor:
        norb = self.norb()
        assert h1e.shape == (norb, norb)
        assert h2e.shape == (norb, norb, norb, norb)
        return self.apply_operator(DenseOperator(norb, h1e, h2e))

    def _apply_array_spatial123(self, h1e: numpy.ndarray, h2e: numpy.ndarray,
                                h3e: numpy.ndarray) -> torch.Tensor:
        """1- through 3-body dense spatial operator (fqe_data.py:1166-1216; the shape of
        reference profiling/profile_3_body.py).  The three-body part is
            - scatter( sum_ij  h3[:, q, i, :, s, j] . gather( gather(C)[i, j] ) )
        i.e. norb^2 gather -> DMMA contraction passes accumulated into one E tensor and one
        scatter, all with the kernels of the 1+2-body path; the lower-rank pieces of
        normal ordering are folded into (h1, h2) on the host first."""
        dev = _require_cuda()
        norb = self.norb()
        h3e = numpy.asarray(h3e)
        if h3e.shape != (norb,) * 6:
            raise ValueError(f"h3e has shape {h3e.shape}, expected {(norb,) * 6}")
        nh1 = numpy.array(h1e, dtype=_C128)
        nh2 = numpy.array(h2e, dtype=_C128)
        # fqe_data.py:1184-1191, vectorised over the loop indices
        nh2 += -numpy.einsum("kjiiab->jkab", h3e) - numpy.einsum("jikiab->jkab", h3e) \
            - numpy.einsum("jkiaib->jkab", h3e)
        nh1 += numpy.einsum("aijijb->ab", h3e)
        out = self._apply_array_spatial12(nh1, nh2)

        lib = _lib.load()
        npair = norb * norb
        ndet = self.lena() * self.lenb()
        align = int(lib.fqeb_gemm_col_align())
        ld = (ndet + align - 1) // align * align
        dvec0 = self.calculate_dvec_spatial()                       # [norb, norb, lena, lenb]
        zero_h1 = numpy.zeros((norb, norb), dtype=_C128)
        acc = torch.zeros((npair + 8, ld), dtype=torch.complex128, device=dev)
        evec = torch.zeros_like(acc)     # padding rows and columns are summed into acc too
        dvec2 = None
        for i in range(norb):
            for j in range(norb):
                m = numpy.ascontiguousarray(h3e[:, :, i, :, :, j].transpose(0, 2, 1, 3),
                                            dtype=_C128)            # [(p, r), (q, s)]
                if not m.any():
                    continue
                op = DenseOperator.from_folded(norb, zero_h1, m, full_pair_space=True)
                rows = int(lib.fqeb_contract_dvec_rows(op.handle, npair))
                if dvec2 is None or dvec2.shape[0] < rows:
                    dvec2 = torch.zeros((rows, ld), dtype=torch.complex128, device=dev)
                _lib.call("fqeb_make_dvec", self._core.handle,
                          self._check_coeff(dvec0[i, j]).data_ptr(), dvec2.data_ptr(), ld, 0,
                          self.lena(), 0, npair, _stream())
                _lib.call("fqeb_contract", op.handle, dvec2.data_ptr(), ld, evec.data_ptr(), ld,
                          ndet, 0, npair, _stream())
                factor = op_factor(op)
                _lib.call("fqeb_zaxpy", acc.numel(), factor.real, factor.imag, evec.data_ptr(),
                          acc.data_ptr(), _stream())
        _lib.call("fqeb_make_coeff", self._core.handle, acc.data_ptr(), ld, 0, self.lena(), -1.0,
                  0.0, out.data_ptr(), _stream())
        return out

    # ---- dvec / coeff (fqe_data.py:2149-2160, 2209-2234, 2309-2334) -----------------
    def calculate_dvec_spatial(self) -> torch.Tensor:
        return self._calculate_dvec_spatial_with_coeff(self.coeff)

    def _calculate_dvec_spatial_with_coeff(self, coeff: torch.Tensor) -> torch.Tensor:
        dev = _require_cuda()
        norb = self.norb()
        dvec = torch.empty((norb, norb, self.lena(), self.lenb()), dtype=torch.complex128,
                           device=dev)
        _lib.call("fqeb_make_dvec", self._core.handle, self._check_coeff(coeff).data_ptr(),
                  dvec.data_ptr(), self.lena() * self.lenb(), 0, self.lena(), 0, norb * norb,
                  _stream())
        return dvec

    def _calculate_coeff_spatial_with_dvec(self, dvec: torch.Tensor) -> torch.Tensor:
        dev = _require_cuda()
        norb = self.norb()
        if tuple(dvec.shape) != (norb, norb, self.lena(), self.lenb()):
            raise ValueError("dvec has the wrong shape")
        out = torch.zeros((self.lena(), self.lenb()), dtype=torch.complex128, device=dev)
        _lib.call("fqeb_make_coeff", self._core.handle, self._check_coeff(dvec).data_ptr(),
                  self.lena() * self.lenb(), 0, self.lena(), 1.0, 0.0, out.data_ptr(), _stream())
        return out

    # ---- diagonal Coulomb (fqe_data.py:263-402) -------------------------------------
    def _dc(self, entry: str, diag, array, inplace: bool) -> torch.Tensor:
        _require_cuda()
        norb = self.norb()
        diag = _host_c128(diag)
        array = _host_c128(array)
        if diag.shape != (norb,) or array.shape != (norb, norb):
            raise ValueError("diagonal Coulomb arrays have the wrong shape")
        data = self.coeff if inplace else self.coeff.clone()
        _lib.call(entry, self._core.handle, diag.ctypes.data, array.ctypes.data,
                  self._check_coeff(data).data_ptr(), _stream())
        return data

    def apply_diagonal_coulomb(self, diag, array, inplace: bool = False) -> torch.Tensor:
        return self._dc("fqeb_dc_apply", diag, array, inplace)

    def evolve_diagonal_coulomb(self, diag, array, inplace: bool = False) -> torch.Tensor:
        return self._dc("fqeb_dc_evolve", diag, array, inplace)

    # ---- one-body diagonal operators (fqe_data.py:153-261) -----------------------------
    def _diag_arrays(self, array) -> Tuple[numpy.ndarray, numpy.ndarray]:
        array = _host_c128(array)
        norb = self.norb()
        if array.size == 2 * norb:
            return numpy.ascontiguousarray(array[:norb]), numpy.ascontiguousarray(array[norb:])
        if array.size != norb:
            raise ValueError('Non-diagonal array passed into a diagonal operator')
        return array, array

    def apply_diagonal_inplace(self, array) -> None:
        """C[a,b] *= sum_{i in a} array[i] + sum_{i in b} array[i]  (fqe_data.py:153-201);
        a 2*norb array holds separate alpha and beta values."""
        _require_cuda()
        aarr, barr = self._diag_arrays(array)
        _lib.call("fqeb_apply_diagonal", self._core.handle, aarr.ctypes.data, barr.ctypes.data,
                  self._check_coeff(self.coeff).data_ptr(), _stream())

    def evolve_diagonal(self, array, inplace: bool = False) -> torch.Tensor:
        """C[a,b] *= exp(sum_{i in a} array[i]) exp(sum_{i in b} array[i])
        (fqe_data.py:203-261); ``array`` is already multiplied by -i*t."""
        _require_cuda()
        aarr, barr = self._diag_arrays(array)
        data = self.coeff if inplace else self.coeff.clone()
        _lib.call("fqeb_evolve_diagonal", self._core.handle, aarr.ctypes.data, barr.ctypes.data,
                  self._check_coeff(data).data_ptr(), _stream())
        return data

    # ---- reduced density matrices (fqe_data.py:1668-1838) ------------------------------------
    _rdm_block_bytes = 1 << 32

    def _rdm_blocks(self, bradata: Optional['FqeData'], want2: bool):
        """sum over alpha-row blocks of  T[ij] = <D_bra[ij] | C>  and  G[ij,kl] =
        <D_bra[ij] | D_ket[kl]>, D = E_ij applied to the state (gather kernel).  Both reductions
        over the determinant index run in one pass of the library's split-K FP64 tensor-core
        kernel (csrc/rdm.cu, ``fqeb_gram_accumulate``): the coefficient block rides along as an
        extra ket row, so T is the last column of the accumulated matrix."""
        dev = _require_cuda()
        bra = self if bradata is None else bradata
        if bra.lena() != self.lena() or bra.lenb() != self.lenb() or bra.norb() != self.norb():
            raise ValueError("bra and ket sectors differ")
        norb, la, lb = self.norb(), self.lena(), self.lenb()
        npair = norb * norb
        # two D blocks of at most _rdm_block_bytes (4 GB) each
        rows = max(1, min(la, self._rdm_block_bytes // max(1, 16 * npair * lb)))
        ncol = npair + 1 if want2 else 1
        acc = torch.zeros((npair, ncol), dtype=torch.complex128, device=dev)
        ket_c, bra_c = self._check_coeff(self.coeff), self._check_coeff(bra.coeff)
        for r0 in range(0, la, rows):
            nr = min(rows, la - r0)
            ld = nr * lb
            dket = None
            if want2 or bra is self:
                dket = torch.empty((npair, ld), dtype=torch.complex128, device=dev)
                _lib.call("fqeb_make_dvec", self._core.handle, ket_c.data_ptr(), dket.data_ptr(),
                          ld, r0, nr, 0, npair, _stream())
            if bra is self:
                dbra = dket
            else:
                dbra = torch.empty((npair, ld), dtype=torch.complex128, device=dev)
                _lib.call("fqeb_make_dvec", self._core.handle, bra_c.data_ptr(), dbra.data_ptr(),
                          ld, r0, nr, 0, npair, _stream())
            _lib.call("fqeb_gram_accumulate", npair, ncol, ld, dbra.data_ptr(), ld,
                      dket.data_ptr() if want2 else None, ld,
                      ket_c.data_ptr() + 16 * r0 * lb, acc.data_ptr(), _stream())
            del dket, dbra
        acc = acc.cpu().numpy()
        return numpy.ascontiguousarray(acc[:, ncol - 1]).reshape(norb, norb), \
            (numpy.ascontiguousarray(acc[:, :npair]).reshape((norb,) * 4) if want2 else None)

    def rdm1(self, bradata: Optional['FqeData'] = None) -> Tuple[numpy.ndarray]:
        """(rdm1,) with rdm1[i,j] = <bra| a+_i a_j |ket>, spin-summed (fqe_data.py:1668-1724)"""
        t1, _ = self._rdm_blocks(bradata, False)
        return (numpy.transpose(t1),)

    def rdm12(self, bradata: Optional['FqeData'] = None) -> Tuple[numpy.ndarray, numpy.ndarray]:
        """(rdm1, rdm2) with rdm2[i,j,k,l] = <bra| a+_i a+_j a_k a_l |ket>
        (fqe_data.py:1726-1807; one algorithm for every filling)."""
        t1, g2 = self._rdm_blocks(bradata, True)
        rdm1 = numpy.transpose(t1)
        rdm2 = -g2.transpose(1, 2, 0, 3)
        for i in range(self.norb()):
            rdm2[:, i, i, :] += rdm1
        return rdm1, numpy.ascontiguousarray(rdm2)

    # ---- individual n-body operators (fqe_data.py:1558-1653, 2385-2580) --------------------
    @staticmethod
    def _op_arrays(dag, undag, norb: int):
        if len(dag) != len(undag):
            raise NotImplementedError("only spin- and number-conserving individual operators "
                                      "(equal numbers of creators and annihilators per spin)")
        for o in list(dag) + list(undag):
            if o < 0 or o >= norb:
                raise ValueError("orbital index out of range")
        return (numpy.ascontiguousarray(dag, dtype=numpy.int32),
                numpy.ascontiguousarray(undag, dtype=numpy.int32))

    def apply_individual_nbody(self, coeff: complex, daga, undaga, dagb, undagb) -> 'FqeData':
        """coeff * prod a+_{daga} prod a_{undaga} prod a+_{dagb} prod a_{undagb} |self>"""
        out = FqeData(self.nalpha(), self.nbeta(), self.norb(), self._core)
        out.coeff.zero_()
        out.apply_individual_nbody_accumulate(coeff, self, daga, undaga, dagb, undagb)
        return out

    def apply_individual_nbody_accumulate(self, coeff: complex, idata: 'FqeData', daga, undaga,
                                          dagb, undagb) -> None:
        """self += coeff * (individual operator) |idata>  (fqe_data.py:1590-1653)"""
        _require_cuda()
        da, ua = self._op_arrays(daga, undaga, self.norb())
        db, ub = self._op_arrays(dagb, undagb, self.norb())
        if idata.coeff.data_ptr() == self.coeff.data_ptr():
            raise ValueError("input and output of an individual n-body apply must differ")
        coeff = complex(coeff)
        _lib.call("fqeb_nbody_accumulate", self._core.handle, coeff.real, coeff.imag,
                  da.ctypes.data, ua.ctypes.data, len(da), db.ctypes.data, ub.ctypes.data, len(db),
                  self._check_coeff(idata.coeff).data_ptr(),
                  self._check_coeff(self.coeff).data_ptr(), _stream())

    def _sparse_scale(self, factor: complex, opa, oha, opb, ohb) -> None:
        def mask(ops):
            m = 0
            for o in ops:
                m |= 1 << int(o)
            return m
        factor = complex(factor)
        _lib.call("fqeb_sparse_scale", self._core.handle, mask(opa), mask(oha), mask(opb),
                  mask(ohb), factor.real, factor.imag, self._check_coeff(self.coeff).data_ptr(),
                  _stream())

    def apply_cos_inplace(self, time: float, ncoeff: complex, opa, oha, opb, ohb) -> None:
        """C *= cos(t |ncoeff|) on the determinants with opa/opb occupied and oha/ohb empty
        (fqe_data.py:2548-2580)"""
        _require_cuda()
        self._sparse_scale(math.cos(time * abs(ncoeff)), opa, oha, opb, ohb)

    def evolve_inplace_individual_nbody_trivial(self, time: float, coeff: complex, opa,
                                                opb) -> None:
        """exp(-i t (T + T^+)) for T = coeff * (product of number operators), in place
        (fqe_data.py:2385-2433); coeff includes the parity due to sorting."""
        _require_cuda()
        n_a, n_b = len(opa), len(opb)
        coeff = complex(coeff) * (-1)**(n_a * (n_a - 1) // 2 + n_b * (n_b - 1) // 2)
        factor = numpy.exp(-time * numpy.real(coeff) * 2.j)
        self._sparse_scale(factor, opa, [], opb, [])

    def evolve_individual_nbody_nontrivial(self, time: float, coeff: complex, daga, undaga, dagb,
                                           undagb) -> 'FqeData':
        """exp(-i t (T + T^+)) |self> for an individual spin-conserving T with T^2 = 0
        (fqe_data.py:2435-2513):  -1 + cos(t sqrt(T T^+)) + cos(t sqrt(T^+ T))
        - i T sin(t sqrt(T^+ T))/sqrt(T^+ T) - i T^+ sin(t sqrt(T T^+))/sqrt(T T^+)."""
        _require_cuda()

        def isolate_number_operators(dag, undag, dagwork, undagwork, number) -> int:
            par = 0
            for current in dag:
                if current in undag:
                    index1 = dagwork.index(current)
                    index2 = undagwork.index(current)
                    par += len(dagwork) - (index1 + 1) + index2
                    dagwork.remove(current)
                    undagwork.remove(current)
                    number.append(current)
            return par

        daga, undaga, dagb, undagb = list(daga), list(undaga), list(dagb), list(undagb)
        dagworka, undagworka = list(daga), list(undaga)
        dagworkb, undagworkb = list(dagb), list(undagb)
        numbera: list = []
        numberb: list = []
        parity = isolate_number_operators(daga, undaga, dagworka, undagworka, numbera)
        parity += isolate_number_operators(dagb, undagb, dagworkb, undagworkb, numberb)
        ncoeff = coeff * (-1)**parity
        absol = numpy.absolute(ncoeff)
        sinfactor = numpy.sin(time * absol) / absol

        out = FqeData(self.nalpha(), self.nbeta(), self.norb(), self._core)
        out.coeff.copy_(self.coeff)
        out.apply_cos_inplace(time, ncoeff, numbera + dagworka, undagworka, numberb + dagworkb,
                              undagworkb)
        out.apply_cos_inplace(time, ncoeff, numbera + undagworka, dagworka, numberb + undagworkb,
                              dagworkb)
        phase = (-1)**((len(daga) + len(undaga)) * (len(dagb) + len(undagb)))
        work_cof = numpy.conj(coeff) * phase * (-1.0j)
        out.apply_individual_nbody_accumulate(work_cof * sinfactor, self, undaga, daga, undagb,
                                              dagb)
        out.apply_individual_nbody_accumulate(coeff * (-1.0j) * sinfactor, self, daga, undaga,
                                              dagb, undagb)
        return out

    # ---- orbital rotation by column operators (fqe_data.py:1476-1535) ---------------------
    def apply_columns_recursive_inplace(self, mat1, mat2) -> None:
        """For icol = 0..norb-1: C <- (1 + sum_i mat1[i,icol] a+_{i,alpha} a_{icol,alpha}) C, then
        the same with mat2 on the beta strings.  Only called from ``Wavefunction.transform``."""
        _require_cuda()
        norb = self.norb()
        mat1, mat2 = _host_c128(mat1), _host_c128(mat2)
        if mat1.shape != (norb, norb) or mat2.shape != (norb, norb):
            raise ValueError("column operators must be norb x norb")
        ptr = self._check_coeff(self.coeff).data_ptr()
        _lib.call("fqeb_apply_columns", self._core.handle, 0, mat1.ctypes.data, ptr, _stream())
        _lib.call("fqeb_apply_columns", self._core.handle, 1, mat2.ctypes.data, ptr, _stream())
