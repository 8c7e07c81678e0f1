"""FciGraph: the determinant addressing system, resident on the GPU.

Host-side mirror of the reference class of the same name
(/root/reference/src/fqe/fci_graph.py:99-570).  The reference builds Python
dicts / numpy tables with C helpers at construction time; here construction
launches two CUDA kernels (csrc/graph.cu) that leave the string tables and the
signed excitation maps in HBM, where the sigma kernels read them.  The host-side
views the reference exposes (``string_alpha_all``, ``alpha_map``, ``_dexca`` ...)
are materialised lazily from the device tables and are bit-identical to the
reference's.
"""
import ctypes
from typing import Dict, Tuple

import numpy as np

from fqe_b200 import lib as _lib


class FciGraph:
    """Knowles-Handy addressing for one (nalpha, nbeta, norb) sector."""

    def __init__(self, nalpha: int, nbeta: int, norb: int) -> None:
        # same argument checks and messages as fci_graph.py:118-132
        if norb < 0:
            raise ValueError(f'norb needs to be >= 0, passed value is {norb}')
        if nalpha < 0:
            raise ValueError(f'nalpha needs to be >= 0, passed value is {nalpha}')
        if nbeta < 0:
            raise ValueError(f'nbeta needs to be >= 0, passed value is {nbeta}')
        if nalpha > norb:
            raise ValueError(f'nalpha needs to be <= norb, passed value is {nalpha}')
        if nbeta > norb:
            raise ValueError(f'nbeta needs to be <= norb, passed value is {nbeta}')
        self._norb = int(norb)
        self._nalpha = int(nalpha)
        self._nbeta = int(nbeta)
        handle = ctypes.c_void_p()
        _lib.call("fqeb_graph_create", self._norb, self._nalpha, self._nbeta,
                  ctypes.byref(handle))
        self._handle = handle
        la, lb = ctypes.c_int64(), ctypes.c_int64()
        _lib.call("fqeb_graph_dims", self._handle, None, None, None, ctypes.byref(la),
                  ctypes.byref(lb))
        self._lena, self._lenb = int(la.value), int(lb.value)
        self._cache: Dict[str, object] = {}

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.load().fqeb_graph_destroy(h)
            except Exception:  # interpreter shutdown
                pass
            self._handle = None

    # ---- handle for the kernels -----------------------------------------------
    @property
    def handle(self) -> ctypes.c_void_p:
        return self._handle

    # ---- sizes ---------------------------------------------------------------------
    def lena(self) -> int:
        return self._lena

    def lenb(self) -> int:
        return self._lenb

    def nalpha(self) -> int:
        return self._nalpha

    def nbeta(self) -> int:
        return self._nbeta

    def norb(self) -> int:
        return self._norb

    # ---- host mirrors -----------------------------------------------------------
    def _strings(self, spin: int) -> np.ndarray:
        key = f"str{spin}"
        if key not in self._cache:
            n = self._lena if spin == 0 else self._lenb
            out = np.zeros(n, dtype=np.uint64)
            _lib.call("fqeb_graph_get_strings", self._handle, spin, out.ctypes.data)
            out.setflags(write=False)
            self._cache[key] = out
        return self._cache[key]

    def _index(self, spin: int) -> Dict[int, int]:
        key = f"ind{spin}"
        if key not in self._cache:
            self._cache[key] = {int(s): i for i, s in enumerate(self._strings(spin))}
        return self._cache[key]

    def z_matrix(self, spin: int) -> np.ndarray:
        """int32[nele, norb] Knowles-Handy Z matrix (fci_graph.py:66-96)."""
        nele = self._nalpha if spin == 0 else self._nbeta
        out = np.zeros((nele, self._norb), dtype=np.int32)
        if out.size:
            _lib.call("fqeb_graph_get_Z", self._handle, spin, out.ctypes.data)
        return out

    def _dense_map(self, spin: int) -> np.ndarray:
        """int32[norb*norb, len]: entry [(i*norb+j), s] = sign*(t+1) for
        a^+_i a_j |s> = sign |t>, 0 if annihilated."""
        key = f"map{spin}"
        if key not in self._cache:
            n = self._lena if spin == 0 else self._lenb
            out = np.zeros((self._norb * self._norb, n), dtype=np.int32)
            if out.size:
                _lib.call("fqeb_graph_get_map", self._handle, spin, out.ctypes.data)
            out.setflags(write=False)
            self._cache[key] = out
        return self._cache[key]

    def _map(self, spin: int, iorb: int, jorb: int) -> np.ndarray:
        row = self._dense_map(spin)[iorb * self._norb + jorb]
        src = np.nonzero(row)[0]
        val = row[src]
        return np.stack([src, np.abs(val) - 1, np.sign(val)], axis=1).astype(np.int32)

    def _dexc(self, spin: int) -> np.ndarray:
        key = f"dexc{spin}"
        if key not in self._cache:
            norb = self._norb
            nele = self._nalpha if spin == 0 else self._nbeta
            n = self._lena if spin == 0 else self._lenb
            lk = nele * (norb - nele + 1)
            dexc = np.zeros((n, lk, 3), dtype=np.int32)
            fill = np.zeros(n, dtype=np.int64)
            for i in range(norb):
                for j in range(norb):
                    m = self._map(spin, i, j)
                    if m.shape[0] == 0:
                        continue
                    tgt = m[:, 1]
                    dexc[tgt, fill[tgt], 0] = m[:, 0]
                    dexc[tgt, fill[tgt], 1] = i * norb + j
                    dexc[tgt, fill[tgt], 2] = m[:, 2]
                    fill[tgt] += 1
            self._cache[key] = dexc
        return self._cache[key]

    # ---- reference API (fci_graph.py:239-399) -----------------------------------
    def alpha_map(self, iorb: int, jorb: int) -> np.ndarray:
        """(source, target, sign) triples of a^+_i a_j on alpha strings."""
        return self._map(0, iorb, jorb)

    def beta_map(self, iorb: int, jorb: int) -> np.ndarray:
        return self._map(1, iorb, jorb)

    @property
    def _alpha_map(self) -> Dict[Tuple[int, int], np.ndarray]:
        return {(i, j): self._map(0, i, j)
                for i in range(self._norb) for j in range(self._norb)}

    @property
    def _beta_map(self) -> Dict[Tuple[int, int], np.ndarray]:
        return {(i, j): self._map(1, i, j)
                for i in range(self._norb) for j in range(self._norb)}

    @property
    def _dexca(self) -> np.ndarray:
        return self._dexc(0)

    @property
    def _dexcb(self) -> np.ndarray:
        return self._dexc(1)

    def string_alpha(self, address: int) -> int:
        return int(self._strings(0)[address])

    def string_beta(self, address: int) -> int:
        return int(self._strings(1)[address])

    def string_alpha_all(self) -> np.ndarray:
        return self._strings(0)

    def string_beta_all(self) -> np.ndarray:
        return self._strings(1)

    def index_alpha(self, address: int) -> int:
        return self._index(0)[address]

    def index_beta(self, address: int) -> int:
        return self._index(1)[address]

    def index_alpha_all(self) -> Dict[int, int]:
        return self._index(0)

    def index_beta_all(self) -> Dict[int, int]:
        return self._index(1)


_GRAPHS: Dict[Tuple[int, int, int, int], FciGraph] = {}


def get_graph(nalpha: int, nbeta: int, norb: int) -> FciGraph:
    """Shared FciGraph per (sector, device): tables are built and uploaded once,
    as FqeData copies share ``_core`` in the reference (fqe_data.py:138-142)."""
    import torch
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    key = (nalpha, nbeta, norb, dev)
    if key not in _GRAPHS:
        _GRAPHS[key] = FciGraph(nalpha, nbeta, norb)
    return _GRAPHS[key]
