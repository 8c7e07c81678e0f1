"""Hamiltonian containers for the hot path (RestrictedHamiltonian, DiagonalCoulomb)."""
from fqe_b200.hamiltonians import hamiltonian, restricted_hamiltonian, diagonal_coulomb  # noqa: F401
