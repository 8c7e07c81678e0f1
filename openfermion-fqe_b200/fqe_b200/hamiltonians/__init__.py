"""Hamiltonian containers for the hot path (RestrictedHamiltonian, DiagonalCoulomb, Diagonal, SparseHamiltonian)."""
from fqe_b200.hamiltonians import (hamiltonian, restricted_hamiltonian, diagonal_coulomb,  # noqa: F401
                                   diagonal_hamiltonian, sparse_hamiltonian)
