"""fqe_b200: B200-native drop-in for OpenFermion-FQE's Hamiltonian-application path.

Public surface mirrors the reference's ``fqe`` package for that path
(/root/reference/src/fqe/__init__.py, _fqe_control.py): ``Wavefunction``,
``get_wavefunction``, ``get_restricted_hamiltonian``,
``get_diagonalcoulomb_hamiltonian``, ``get_diagonal_hamiltonian``, ``apply``, ``time_evolve``,
``apply_generated_unitary``, ``vdot``, ``dot``, ``expectationValue``.
Importing the package does not need a GPU; every compute call does (there is no
CPU fallback) and raises ``fqe_b200.lib.FqeB200Error`` otherwise.
"""
from fqe_b200 import settings  # noqa: F401
from fqe_b200._fqe_control import (Wavefunction, apply, apply_generated_unitary, dot,
                                   expectationValue, get_diagonal_hamiltonian,
                                   get_diagonalcoulomb_hamiltonian,
                                   get_restricted_hamiltonian, get_sparse_hamiltonian,
                                   get_wavefunction, time_evolve,
                                   vdot)

__version__ = "0.1.0"
__all__ = [
    "Wavefunction", "apply", "apply_generated_unitary", "dot", "expectationValue",
    "get_diagonal_hamiltonian", "get_diagonalcoulomb_hamiltonian", "get_restricted_hamiltonian", "get_sparse_hamiltonian", "get_wavefunction",
    "time_evolve", "vdot", "settings"
]
