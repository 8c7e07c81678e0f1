"""RestrictedHamiltonian: identical alpha/beta tensors over spatial orbitals.

API-compatible with
/root/reference/src/fqe/hamiltonians/restricted_hamiltonian.py:23-155.
H = e_0 + sum h1[i,j] a+_i a_j + sum h2[i,j,k,l] a+_i a+_j a_k a_l + ...
"""
from typing import Dict, Tuple

import numpy

from fqe_b200.hamiltonians import hamiltonian


def _same_tensors(lhs: Dict[int, numpy.ndarray], rhs: Dict[int, numpy.ndarray]) -> bool:
    if lhs.keys() != rhs.keys():
        return False
    return all(lhs[k].shape == rhs[k].shape and numpy.allclose(lhs[k], rhs[k]) for k in lhs)


class RestrictedHamiltonian(hamiltonian.Hamiltonian):
    """Dense spin-free operator given as a tuple of 1..4-body tensors."""

    def __init__(self, tensors: Tuple[numpy.ndarray, ...], e_0: complex = 0.0 + 0.0j) -> None:
        super().__init__(e_0=e_0)
        self._tensor: Dict[int, numpy.ndarray] = {}
        for nbody, matrix in enumerate(tensors, start=1):
            if not isinstance(matrix, numpy.ndarray):
                raise TypeError("tensors should be a tuple of numpy.ndarray")
            if matrix.ndim % 2:
                raise ValueError("input tensor has an odd rank")
            self._tensor[2 * nbody] = matrix
        assert self._tensor, "No matrix elements passed into the RestrictedHamiltonian."
        self._quadratic = len(self._tensor) == 1 and 2 in self._tensor
        self._dim = next(iter(self._tensor.values())).shape[0]

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, RestrictedHamiltonian):
            return NotImplemented
        return self.e_0() == other.e_0() and _same_tensors(self._tensor, other._tensor)

    def dim(self) -> int:
        return self._dim

    def rank(self) -> int:
        return 2 * len(self._tensor)

    def tensor(self, rank: int) -> numpy.ndarray:
        return self._tensor[rank]

    def tensors(self) -> Tuple[numpy.ndarray, ...]:
        return tuple(self._tensor[2 * (k + 1)] for k in range(len(self._tensor)))

    def quadratic(self) -> bool:
        return self._quadratic

    def iht(self, time: float) -> Tuple[numpy.ndarray, ...]:
        return tuple(-1.0j * time * t for t in self.tensors())

    def calc_diag_transform(self) -> numpy.ndarray:
        _, trans = numpy.linalg.eigh(self._tensor[2])
        return trans

    def transform(self, trans: numpy.ndarray) -> numpy.ndarray:
        return trans.conj().T @ self._tensor[2] @ trans
