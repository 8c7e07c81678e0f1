"""Diagonal: one-body Hamiltonian that is diagonal in the determinant basis,
H = e_0 + sum_r h[r] a+_r a_r.

API-compatible with /root/reference/src/fqe/hamiltonians/diagonal_hamiltonian.py:26-121.
It is what the quadratic branch of ``Wavefunction.time_evolve`` evolves between the two
orbital rotations.
"""
import copy

import numpy

from fqe_b200.hamiltonians import hamiltonian


class Diagonal(hamiltonian.Hamiltonian):

    def __init__(self, hdiag: numpy.ndarray, e_0: complex = 0.0 + 0.0j) -> None:
        super().__init__(e_0=e_0)
        if hdiag.ndim != 1:
            raise ValueError(
                "Incorrect dimension passed for DiagonalHamiltonian elements. "
                f"Must have hdiag.ndim = 1 but hdiag.ndim = {hdiag.ndim}.")
        self._hdiag = hdiag
        self._dim = self._hdiag.shape[0]

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, Diagonal):
            return NotImplemented
        return self.e_0() == other.e_0() and all(self._hdiag == other._hdiag)

    def dim(self) -> int:
        return self._dim

    def rank(self) -> int:
        return 2

    def diagonal(self) -> bool:
        return True

    def quadratic(self) -> bool:
        return True

    def diag_values(self) -> numpy.ndarray:
        return self._hdiag

    def iht(self, time: float) -> 'Diagonal':
        out = copy.deepcopy(self)
        out._hdiag = out._hdiag * (-1.0j * time)
        return out
