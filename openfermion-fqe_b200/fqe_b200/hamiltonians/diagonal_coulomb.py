"""DiagonalCoulomb: H = e_0 + sum_r f_r n_r + sum_rs v_rs n_r n_s.

API-compatible with /root/reference/src/fqe/hamiltonians/diagonal_coulomb.py:23-123.
Accepts either the rank-2 v_rs array or a rank-4 two-body tensor in the dense
Hamiltonian layout, from which diag[k] = h[k,k,k,k] and v[i,j] = -h[i,j,i,j].
"""
from typing import Dict, Tuple

import numpy

from fqe_b200.hamiltonians import hamiltonian
from fqe_b200.hamiltonians.restricted_hamiltonian import _same_tensors


class DiagonalCoulomb(hamiltonian.Hamiltonian):
    """Two-body operator that is diagonal in the determinant basis."""

    def __init__(self, h2e: numpy.ndarray, e_0: complex = 0.0 + 0.0j) -> None:
        super().__init__(e_0=e_0)
        self._dim = h2e.shape[0]
        self._tensor: Dict[int, numpy.ndarray] = {}
        if h2e.ndim == 2:
            self._tensor[1] = numpy.zeros(self._dim, dtype=h2e.dtype)
            self._tensor[2] = h2e
        elif h2e.ndim == 4:
            idx = numpy.arange(self._dim)
            self._tensor[1] = h2e[idx, idx, idx, idx].copy()
            ii, jj = numpy.meshgrid(idx, idx, indexing="ij")
            self._tensor[2] = -h2e[ii, jj, ii, jj]
        else:
            raise ValueError("DiagonalCoulomb expects a rank-2 or rank-4 array")

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, DiagonalCoulomb):
            return NotImplemented
        return self.e_0() == other.e_0() and _same_tensors(self._tensor, other._tensor)

    def dim(self) -> int:
        return self._dim

    def diagonal_coulomb(self) -> bool:
        return True

    def rank(self) -> int:
        return 4

    def iht(self, time: float) -> Tuple[numpy.ndarray, ...]:
        return tuple(-1.0j * time * self._tensor[k] for k in (1, 2))
