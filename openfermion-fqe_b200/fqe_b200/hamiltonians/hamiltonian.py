"""Base class for Hamiltonian containers.

API-compatible with /root/reference/src/fqe/hamiltonians/hamiltonian.py:26-135:
a passive holder of tensors plus type predicates that ``Wavefunction.apply`` /
``time_evolve`` dispatch on.
"""
from abc import ABCMeta, abstractmethod
from typing import Any, Tuple

import numpy


class Hamiltonian(metaclass=ABCMeta):
    """Common interface: scalar part ``e_0`` and structural predicates."""

    def __init__(self, e_0: complex = 0.0 + 0.0j):
        self._conserve_number = True
        self._e_0 = e_0

    @abstractmethod
    def dim(self) -> int:
        """Orbital dimension of the tensors."""

    @abstractmethod
    def rank(self) -> int:
        """Rank of the largest tensor."""

    def quadratic(self) -> bool:
        return False

    def diagonal(self) -> bool:
        return False

    def diagonal_coulomb(self) -> bool:
        return False

    def conserve_number(self) -> bool:
        return self._conserve_number

    def e_0(self):
        return self._e_0

    def iht(self, time: float) -> Any:
        """Tensors premultiplied by -i*time."""
        return tuple()

    def tensors(self) -> Tuple[numpy.ndarray, ...]:
        return tuple()

    def diag_values(self) -> numpy.ndarray:
        return numpy.empty(0)

    def calc_diag_transform(self) -> numpy.ndarray:
        return numpy.empty(0)

    def transform(self, trans: numpy.ndarray) -> numpy.ndarray:
        return numpy.empty(0)
