"""Module-level API, forwarding to Wavefunction as the reference's facade does
(/root/reference/src/fqe/_fqe_control.py:60-96, 145-198, 356-371, 431-462, 495-575)."""
from typing import List, Optional, Tuple

import numpy

from fqe_b200 import wavefunction as _wfn
from fqe_b200.hamiltonians import (diagonal_coulomb, diagonal_hamiltonian, restricted_hamiltonian,
                                   sparse_hamiltonian)


def Wavefunction(param: List[List[int]], broken=None) -> '_wfn.Wavefunction':
    """Wavefunction([[n_electrons, 2*s_z, norb], ...])  (_fqe_control.py:145-159)."""
    return _wfn.Wavefunction(param, broken=broken)


def get_wavefunction(nele: int, m_s: int, norb: int) -> '_wfn.Wavefunction':
    """Single-sector wavefunction (_fqe_control.py:162-177)."""
    return _wfn.Wavefunction([[nele, m_s, norb]])


def get_restricted_hamiltonian(tensors: Tuple[numpy.ndarray, ...], e_0: complex = 0. + 0.j):
    """(_fqe_control.py:559-574)"""
    return restricted_hamiltonian.RestrictedHamiltonian(tensors, e_0=e_0)


def get_diagonalcoulomb_hamiltonian(h2e: numpy.ndarray, e_0: complex = 0. + 0.j):
    """(_fqe_control.py:495-508)"""
    return diagonal_coulomb.DiagonalCoulomb(h2e, e_0=e_0)


def get_diagonal_hamiltonian(hdiag: numpy.ndarray, e_0: complex = 0. + 0.j):
    """(_fqe_control.py:511-523)"""
    return diagonal_hamiltonian.Diagonal(hdiag, e_0=e_0)


def get_sparse_hamiltonian(operators, conserve_spin: bool = True, e_0: complex = 0. + 0.j):
    """(_fqe_control.py:577-597); ``operators``: FermionOperator-like, terms mapping or string"""
    return sparse_hamiltonian.SparseHamiltonian(operators, conserve_spin=conserve_spin, e_0=e_0)


def apply(ops, wfn: '_wfn.Wavefunction') -> '_wfn.Wavefunction':
    """H|wfn>  (_fqe_control.py:356-371)."""
    return wfn.apply(ops)


def time_evolve(wfn: '_wfn.Wavefunction', time: float, hamil,
                inplace: bool = False) -> '_wfn.Wavefunction':
    """exp(-i t H)|wfn>  (_fqe_control.py:180-198)."""
    return wfn.time_evolve(time, hamil, inplace)


def apply_generated_unitary(wfn: '_wfn.Wavefunction',
                            time: float,
                            algo: str,
                            hamil,
                            accuracy: float = 1.0E-15,
                            expansion: int = 30,
                            spec_lim: Optional[List[float]] = None) -> '_wfn.Wavefunction':
    """Polynomial propagator  (_fqe_control.py:60-96)."""
    return wfn.apply_generated_unitary(time, algo, hamil, accuracy=accuracy, expansion=expansion,
                                       spec_lim=spec_lim)


def vdot(wfn1: '_wfn.Wavefunction', wfn2: '_wfn.Wavefunction') -> complex:
    """<wfn1|wfn2> with conjugation  (_fqe_control.py:448-462)."""
    return wfn1.vdot(wfn2)


def dot(wfn1: '_wfn.Wavefunction', wfn2: '_wfn.Wavefunction') -> complex:
    """sum wfn1*wfn2 without conjugation  (_fqe_control.py:431-445)."""
    total = 0.0 + 0.0j
    for key in wfn1.sectors():
        total += complex((wfn1.get_coeff_device(key) * wfn2.get_coeff_device(key)).sum().item())
    return total


def expectationValue(wfn: '_wfn.Wavefunction', ops, brawfn=None) -> complex:
    """(_fqe_control.py:374-392)"""
    return wfn.expectationValue(ops, brawfn)
