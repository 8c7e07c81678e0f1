"""Wavefunction: dictionary of (n, sz) sectors living on the GPU.

Host-side mirror of the reference class (/root/reference/src/fqe/wavefunction.py)
for the Hamiltonian-application path: ``apply``, ``apply_generated_unitary``
(Taylor and Chebyshev), ``time_evolve`` and the BLAS-1 helpers keep the
reference's signatures and semantics - including its quirks (the e_0 phase is
applied twice on the Taylor branch of ``time_evolve``, wavefunction.py:604-605 and
1051-1052; the Taylor series is capped at ``expansion`` terms and raises
``RuntimeError``) - while every coefficient stays in HBM and every operation is a
CUDA kernel of libfqe_b200.so.

Only number- and spin-conserving wavefunctions and the Hamiltonian classes of the
hot path (RestrictedHamiltonian with 1- and 2-body terms, DiagonalCoulomb) are
handled; anything else raises ``NotImplementedError`` rather than falling back.
"""
import copy
import math
import os
import pickle
from typing import Dict, KeysView, List, Optional, Tuple, Union

import numpy
import torch
from scipy import linalg
from scipy.special import factorial, jv

from fqe_b200.fqe_data import DenseOperator, FqeData
from fqe_b200.hamiltonians import (diagonal_coulomb, diagonal_hamiltonian, hamiltonian,
                                   restricted_hamiltonian, sparse_hamiltonian)


def alpha_beta_electrons(nele: int, m_s: int) -> Tuple[int, int]:
    """(nalpha, nbeta) from particle number and 2*S_z (util.py:29-50)."""
    if nele < 0:
        raise ValueError('Cannot have negative electrons')
    if nele < abs(m_s):
        raise ValueError('Spin quantum number exceeds physical limits')
    if (nele + m_s) % 2 != 0:
        raise ValueError('Parity of spin quantum number and number of electrons is incompatible')
    nalpha = int(nele + m_s) // 2
    return nalpha, nele - nalpha


def build_hamiltonian(ops, norb: int = 0, e_0: complex = 0.0 + 0.0j) -> hamiltonian.Hamiltonian:
    """Tuple -> RestrictedHamiltonian, Hamiltonian -> itself
    (fqe_decorators.py:45-73; the FermionOperator branch needs openfermion and is
    outside this path)."""
    if isinstance(ops, hamiltonian.Hamiltonian):
        return ops
    if isinstance(ops, tuple):
        for t in ops:
            if not isinstance(t, numpy.ndarray):
                raise TypeError("Arguments in tuple should be numpy.array")
        if norb != 0 and ops[0].shape[0] == norb:
            return restricted_hamiltonian.RestrictedHamiltonian(ops, e_0=e_0)
        raise NotImplementedError("only spatial-orbital (restricted) tensors are supported")
    raise TypeError('Expected a Hamiltonian or a tuple of numpy arrays but received {}.'.format(
        type(ops)))


def rotation_factors(rot: numpy.ndarray, low: Optional[numpy.ndarray] = None,
                     upp: Optional[numpy.ndarray] = None):
    """Host side of ``Wavefunction.transform`` (wavefunction.py:838-909): (P, L, U, M) with
    rot^H = P L U from ``scipy.linalg.lu`` and M the norb x norb matrix of column operators,
    M = inv(U') - strict_lower(L') - 1, where L', U' are the Hermitian conjugates of the factors
    with the diagonal moved into the upper one.  With external ``low``, ``upp`` (rot == low @
    upp, the way back in the quadratic time evolution) they are used as they are."""

    def transpose_matrix(low, upp):
        ndim = low.shape[0]
        lowt, uppt = low.astype(numpy.complex128), upp.astype(numpy.complex128)
        for irow in range(ndim):
            uppt[irow, irow + 1:] /= uppt[irow, irow]
            lowt[irow, irow], uppt[irow, irow] = uppt[irow, irow], lowt[irow, irow]
            for icol in range(irow):
                lowt[irow, icol] *= lowt[icol, icol]
        return uppt.T.conj(), lowt.T.conj()

    def process_matrix(low, upp):
        ndim = low.shape[0]
        output = linalg.solve_triangular(upp, numpy.identity(ndim))
        return output - numpy.tril(low, -1) - numpy.identity(ndim)

    if low is None:
        perm, low, upp = linalg.lu(rot.transpose().conjugate())
        lowt, uppt = transpose_matrix(low, upp)
        return perm, low, upp, process_matrix(lowt, uppt)
    return None, low, upp, process_matrix(low, upp)


class Wavefunction:
    """A state vector as a set of FqeData sectors keyed by (nele, m_s)."""

    def __init__(self,
                 param: Optional[List[List[int]]] = None,
                 broken: Optional[Union[List[str], str]] = None) -> None:
        if broken:
            raise NotImplementedError(
                "symmetry-broken wavefunctions are outside the B200 hot path")
        self._conserve_spin = True
        self._conserve_number = True
        self._conserved: Dict[str, int] = {}
        self._norb = 0
        self._civec: Dict[Tuple[int, int], FqeData] = {}
        if param:
            norbs = set(x[2] for x in param)
            if len(norbs) != 1:
                raise ValueError('Number of orbitals is not consistent')
            self._norb = list(norbs)[0]
            for nele, m_s, _ in param:
                nalpha, nbeta = alpha_beta_electrons(nele, m_s)
                self._civec[(nele, m_s)] = FqeData(nalpha, nbeta, self._norb)
            self._conserved['n'] = param[0][0]
            self._conserved['s_z'] = param[0][1]

    # ---- container protocol ----------------------------------------------------------
    def sector(self, key: Tuple[int, int]) -> FqeData:
        return self._civec[key]

    def sectors(self) -> KeysView[Tuple[int, int]]:
        return self._civec.keys()

    def conserve_number(self) -> bool:
        return self._conserve_number

    def conserve_spin(self) -> bool:
        return self._conserve_spin

    def norb(self) -> int:
        return self._norb

    def get_coeff(self, key: Tuple[int, int]) -> numpy.ndarray:
        """Host copy (numpy complex128) of a sector, as the reference returns numpy
        (wavefunction.py:613-624)."""
        return self._civec[key].to_numpy()

    def get_coeff_device(self, key: Tuple[int, int]) -> torch.Tensor:
        """The resident CUDA tensor of a sector (no copy)."""
        return self._civec[key].coeff

    def __getitem__(self, key: Tuple[int, int]) -> complex:
        astr, bstr = key
        na, nb = bin(astr).count("1"), bin(bstr).count("1")
        return self._civec[(na + nb, na - nb)][key]

    def __setitem__(self, key: Tuple[int, int], value: complex) -> None:
        astr, bstr = key
        na, nb = bin(astr).count("1"), bin(bstr).count("1")
        self._civec[(na + nb, na - nb)][key] = value

    def __deepcopy__(self, memodict={}) -> 'Wavefunction':
        out = self.empty_copy(zero=False)
        for key, sec in self._civec.items():
            out._civec[key].coeff.copy_(sec.coeff)
        return out

    def empty_copy(self, zero: bool = True) -> 'Wavefunction':
        out = Wavefunction()
        out._norb = self._norb
        out._conserved = dict(self._conserved)
        for key, sec in self._civec.items():
            out._civec[key] = sec.empty_copy(zero=zero)
        return out

    def set_wfn(self, strategy: str = 'ones', raw_data=None) -> None:
        if strategy == 'from_data' and not raw_data:
            raise ValueError('No data provided for set_wfn')
        if strategy == 'from_data':
            for key, data in raw_data.items():
                self._civec[key].set_wfn(strategy='from_data', raw_data=data)
        elif strategy == 'hartree-fock':
            if len(self.sectors()) != 1:
                raise ValueError("Hartree-Fock wf initialization only works "
                                 "with single sector wavefunctions")
            for sec in self._civec.values():
                sec.set_wfn(strategy=strategy)
        else:
            if strategy == 'zeros':
                strategy = 'zero'
            for sec in self._civec.values():
                sec.set_wfn(strategy=strategy)
        if strategy == 'random':
            self.normalize()

    # ---- BLAS-1 (wavefunction.py:241-254, 296-312, 767-777) -------------------------
    def ax_plus_y(self, sval: complex, wfn: 'Wavefunction') -> None:
        if self._civec.keys() != wfn._civec.keys():
            raise ValueError('inconsistent sectors in Wavefunction.ax_plus_y')
        for key in self._civec:
            self._civec[key].ax_plus_y(sval, wfn._civec[key])

    def __add__(self, other: 'Wavefunction') -> 'Wavefunction':
        out = copy.deepcopy(self)
        out.ax_plus_y(1.0, other)
        return out

    def __iadd__(self, other: 'Wavefunction') -> 'Wavefunction':
        self.ax_plus_y(1.0, other)
        return self

    def __sub__(self, other: 'Wavefunction') -> 'Wavefunction':
        out = copy.deepcopy(self)
        out.ax_plus_y(-1.0, other)
        return out

    def norm(self) -> float:
        return math.sqrt(sum(sec.norm()**2 for sec in self._civec.values()))

    def normalize(self) -> None:
        self.scale(1.0 / self.norm())

    def scale(self, sval: complex) -> None:
        sval = complex(sval)
        for sec in self._civec.values():
            sec.scale(sval)

    def vdot(self, other: 'Wavefunction') -> complex:
        """<self|other> (util.vdot / fqe.vdot, _fqe_control.py:448-462)."""
        if self._civec.keys() != other._civec.keys():
            raise ValueError('inconsistent sectors in vdot')
        return sum(self._civec[k].vdot(other._civec[k]) for k in self._civec)

    def _axpy_norm(self, sval: complex, work: 'Wavefunction') -> float:
        return math.sqrt(
            sum(self._civec[k].axpy_norm(sval, work._civec[k])**2 for k in self._civec))

    # ---- apply (wavefunction.py:314-440) ------------------------------------------------
    def apply(self, hamil) -> 'Wavefunction':
        """Return H|psi> for a Hamiltonian object or a tuple of dense tensors."""
        hamil = build_hamiltonian(hamil, norb=self.norb())
        if not hamil.conserve_number():
            raise TypeError('Number non-conserving hamiltonian passed to'
                            ' number conserving wavefunction')
        if isinstance(hamil, sparse_hamiltonian.SparseHamiltonian):
            return self._apply_few_nbody(hamil)
        if isinstance(hamil, diagonal_hamiltonian.Diagonal):
            return self._apply_diagonal(hamil)
        if isinstance(hamil, diagonal_coulomb.DiagonalCoulomb):
            return self._apply_diagonal_coulomb(hamil)
        if isinstance(hamil, restricted_hamiltonian.RestrictedHamiltonian):
            if hamil.dim() != self._norb:
                raise ValueError('Hamiltonian has incorrect size: expected {} provided {}'.format(
                    self._norb, hamil.dim()))
            return self._apply_array(hamil.tensors(), hamil.e_0())
        raise NotImplementedError(
            f"{type(hamil).__name__} is outside the B200 hot path (RestrictedHamiltonian, "
            "DiagonalCoulomb, Diagonal, SparseHamiltonian)")

    def _dense_operator(self, array: Tuple[numpy.ndarray, ...]) -> DenseOperator:
        if len(array) < 1 or len(array) > 4:
            raise ValueError("Number of operators in tuple must be between 1 and 4.")
        if len(array) > 2:
            raise NotImplementedError("prepared operators cover 1- and 2-body tensors; "
                                      "3-body tuples go through FqeData.apply")
        if array[0].shape[0] != self._norb:
            raise NotImplementedError("only spatial-orbital (restricted) tensors are supported")
        return DenseOperator(self._norb, array[0], array[1] if len(array) == 2 else None)

    def _apply_operator(self, op: DenseOperator, e_0: complex = 0.0) -> 'Wavefunction':
        out = self.empty_copy(zero=False)
        for key, sec in self._civec.items():
            out._civec[key].coeff = sec.apply_operator(op)
        if numpy.abs(e_0) > 1.e-15:
            out.ax_plus_y(e_0, self)
        return out

    def _apply_array(self, array: Tuple[numpy.ndarray, ...], e_0: complex) -> 'Wavefunction':
        if len(array) == 3:
            if array[0].shape[0] != self._norb:
                raise NotImplementedError("only spatial-orbital (restricted) tensors are supported")
            out = self.empty_copy(zero=False)
            for key, sec in self._civec.items():
                out._civec[key].coeff = sec._apply_array_spatial123(array[0], array[1], array[2])
            if numpy.abs(e_0) > 1.e-15:
                out.ax_plus_y(e_0, self)
            return out
        return self._apply_operator(self._dense_operator(array), e_0)

    def _apply_diagonal(self, hamil: diagonal_hamiltonian.Diagonal) -> 'Wavefunction':
        """(wavefunction.py:401-419)"""
        out = copy.deepcopy(self)
        for sec in out._civec.values():
            sec.apply_diagonal_inplace(hamil.diag_values())
        if numpy.abs(hamil.e_0()) > 1.e-15:
            out.ax_plus_y(hamil.e_0(), self)
        return out

    def _apply_diagonal_coulomb(self, hamil: diagonal_coulomb.DiagonalCoulomb) -> 'Wavefunction':
        out = copy.deepcopy(self)
        diag, array = hamil._tensor[1], hamil._tensor[2]
        for sec in out._civec.values():
            sec.apply_diagonal_coulomb(diag, array, inplace=True)
        if numpy.abs(hamil.e_0()) > 1.e-15:
            out.ax_plus_y(hamil.e_0(), self)
        return out

    # ---- polynomial propagators (wavefunction.py:509-611) ----------------------------
    def apply_generated_unitary(self,
                                time: float,
                                algo: str,
                                hamil,
                                accuracy: float = 1.0E-15,
                                expansion: int = 30,
                                spec_lim: Optional[List[float]] = None) -> 'Wavefunction':
        hamil = build_hamiltonian(hamil, norb=self.norb())
        assert isinstance(hamil, hamiltonian.Hamiltonian)
        if not isinstance(expansion, int):
            raise TypeError("expansion must be an int. You provided {}".format(expansion))
        assert algo in ['taylor', 'chebyshev']
        sparse = isinstance(hamil, sparse_hamiltonian.SparseHamiltonian)
        if not sparse and not isinstance(hamil, restricted_hamiltonian.RestrictedHamiltonian):
            raise NotImplementedError(
                "apply_generated_unitary covers RestrictedHamiltonian and SparseHamiltonian")
        base = self
        self.last_expansion_order = 0

        if algo == 'taylor' and sparse:
            # term-by-term application of the prepared operator (wavefunction.py:550-561)
            ham_iht = hamil.iht(time)
            time_evol = copy.deepcopy(base)
            work = copy.deepcopy(base)
            for order in range(1, expansion):
                work = work.apply(ham_iht)
                coeff = 1.0 / factorial(order)
                wnorm = time_evol._axpy_norm(coeff, work)
                if wnorm * numpy.abs(coeff) < accuracy:
                    break
            else:
                raise RuntimeError("maximum taylor expansion limit reached")
            self.last_expansion_order = order

        elif algo == 'taylor':
            # -i*t*H is purely imaginary for real integrals: the operator is prepared
            # once (real-GEMM mode) and reused by every term.  The tuple re-wrap of the
            # reference drops e_0 inside the loop (fqe_decorators.py:68-73).
            ham_arrays = hamil.iht(time)
            op = self._dense_operator(ham_arrays) if len(ham_arrays) <= 2 else None
            time_evol = copy.deepcopy(base)
            if op is not None and len(self._civec) == 1:
                # one sector: the whole recurrence runs inside the library (fqeb_taylor)
                (sec,) = time_evol._civec.values()
                order = sec.taylor_inplace(op, accuracy, expansion)
                self.last_expansion_order = order
                if numpy.abs(hamil.e_0() * time) > 1.e-15:
                    time_evol.scale(numpy.exp(-1.j * time * hamil.e_0()))
                return time_evol
            work = copy.deepcopy(base)
            for order in range(1, expansion):
                work = work._apply_operator(op) if op is not None else \
                    work._apply_array(ham_arrays, 0.0)
                coeff = 1.0 / factorial(order)
                wnorm = time_evol._axpy_norm(coeff, work)
                if wnorm * numpy.abs(coeff) < accuracy:
                    break
            else:
                raise RuntimeError("maximum taylor expansion limit reached")
            self.last_expansion_order = order

        else:
            assert spec_lim, 'Spectral range was not provided. Provide upper and lower limits.'
            wprime = 0.9875
            ascale = (spec_lim[1] - spec_lim[0]) / (2.0 * wprime)
            eshift = -(spec_lim[0] + ascale * wprime)
            tensors = hamil.tensors()
            e_0 = hamil.e_0()
            op = self._dense_operator(tensors) if not sparse and len(tensors) <= 2 else None

            def _h(x: 'Wavefunction') -> 'Wavefunction':
                if sparse:
                    return x.apply(hamil)
                return x._apply_operator(op, e_0) if op is not None else x._apply_array(tensors, e_0)

            time_evol = copy.deepcopy(base)
            time_evol.scale(jv(0, ascale * time))
            minus = copy.deepcopy(base)
            current = _h(minus)
            current.ax_plus_y(eshift, minus)
            current.scale(1.0 / ascale)
            time_evol.ax_plus_y(2.0 * jv(1, ascale * time) * (-1.j), current)
            for order in range(2, expansion):
                minus.scale(-1.0)
                minus.ax_plus_y(2.0 / ascale, _h(current))
                minus.ax_plus_y(2.0 * eshift / ascale, current)
                current, minus = minus, current
                coeff = 2.0 * jv(order, ascale * time) * (-1.j)**order
                time_evol.ax_plus_y(coeff, current)
                if current.norm() * numpy.abs(coeff) < accuracy:
                    break
            else:
                raise RuntimeError("maximum chebyshev expansion limit reached")
            self.last_expansion_order = order
            time_evol.scale(numpy.exp(eshift * time * 1.j))

        if numpy.abs(hamil.e_0() * time) > 1.e-15:
            time_evol.scale(numpy.exp(-1.j * time * hamil.e_0()))
        return time_evol

    # ---- time evolution (wavefunction.py:961-1098) --------------------------------------
    def time_evolve(self, time: float, hamil, inplace: bool = False) -> 'Wavefunction':
        hamil = build_hamiltonian(hamil, norb=self.norb())
        assert isinstance(hamil, hamiltonian.Hamiltonian)
        if not hamil.conserve_number():
            raise TypeError('Number non-conserving hamiltonian passed to'
                            ' number conserving wavefunction')
        individual = isinstance(hamil, sparse_hamiltonian.SparseHamiltonian) and \
            hamil.is_individual()
        work_wfn = self
        if not individual:
            is_diag = (hamil.quadratic() and hamil.diagonal()) or hamil.diagonal_coulomb()
            if inplace and (not is_diag and not hamil.quadratic()):
                raise ValueError("Inplace is not implemented for this case")
            work_wfn = self if inplace else copy.deepcopy(self)

        if individual:
            final_wfn = self._evolve_individual_nbody(time, hamil, inplace)
        elif isinstance(hamil, sparse_hamiltonian.SparseHamiltonian):
            final_wfn = work_wfn.apply_generated_unitary(time, 'taylor', hamil)
        elif hamil.quadratic():
            # (wavefunction.py:1013-1034) diagonal: one pass; otherwise rotate the orbitals to
            # the eigenbasis of h1, evolve the diagonal, rotate back with the same L, U factors
            if hamil.dim() != self._norb:
                raise NotImplementedError("only spatial-orbital (restricted) tensors are supported")
            if hamil.diagonal():
                ihtdiag = -1.j * time * hamil.diag_values()
                final_wfn = work_wfn._evolve_diagonal(ihtdiag, inplace)
            else:
                transformation = hamil.calc_diag_transform()
                permu, low, upp, work_wfn = work_wfn.transform(transformation)
                ci_trans = transformation @ permu
                h1e = hamil.transform(ci_trans)
                ihtdiag = -1.j * time * h1e.diagonal()
                evolved_wfn = work_wfn._evolve_diagonal(ihtdiag, inplace=True)
                _, _, _, final_wfn = evolved_wfn.transform(ci_trans.T.conj(), low, upp)
        elif hamil.diagonal_coulomb():
            diag, vij = hamil.iht(time)
            final_wfn = work_wfn._evolve_diagonal_coulomb_inplace(diag, vij)
        elif isinstance(hamil, restricted_hamiltonian.RestrictedHamiltonian):
            final_wfn = work_wfn.apply_generated_unitary(time, 'taylor', hamil)
        else:
            raise NotImplementedError(
                f"time_evolve with {type(hamil).__name__} is outside the B200 hot path")

        if numpy.abs(hamil.e_0()) > 1.0e-15:
            final_wfn.scale(numpy.exp(-1.j * time * hamil.e_0()))
        self.last_expansion_order = getattr(work_wfn, "last_expansion_order", 0)
        return final_wfn

    # ---- RDMs, save / read (wavefunction.py:1357-1415, 726-765) -------------------------------
    def _compute_rdm(self, rank: int, brawfn: Optional['Wavefunction'] = None):
        """Tuple of spin-summed RDMs up to ``rank`` (1 or 2), summed over the sectors"""
        assert 0 < rank < 5
        if rank > 2:
            raise NotImplementedError("3- and 4-particle RDMs are outside the B200 hot path")
        out = None
        for key, sector in self._civec.items():
            assert brawfn is None or key in brawfn.sectors()
            bra = None if brawfn is None else brawfn._civec[key]
            tmp = sector.rdm1(bra) if rank == 1 else sector.rdm12(bra)
            out = tmp if out is None else tuple(a + b for a, b in zip(out, tmp))
        return out

    def save(self, filename: str, path: str = os.getcwd()) -> None:
        """Write path/filename in the reference's own file layout
        (wavefunction.py:743-765: [symmetry_map, conserved, conserve_spin, conserve_number, norb,
        [key, FqeData]...]); the sector entries re-create ``fqe.fqe_data.FqeData`` objects when
        the file is read where the reference is installed (fqe_b200/wfn_io.py)."""
        from fqe_b200 import wfn_io
        sectors = {key: (sec.nalpha(), sec.nbeta(), sec.to_numpy())
                   for key, sec in self._civec.items()}
        with open(os.path.join(path, filename), 'w+b') as fh:
            wfn_io.dump(fh, self._conserved, self._norb, sectors, self._conserve_spin,
                        self._conserve_number)

    def read(self, filename: str, path: str = os.getcwd()) -> None:
        """Initialise from a file written by this package OR by the reference's
        ``Wavefunction.save`` (wavefunction.py:726-741).  Nothing but numpy arrays is ever
        un-pickled: classes of the reference package are mapped to attribute bags."""
        from fqe_b200 import wfn_io
        with open(os.path.join(path, filename), 'r+b') as fh:
            data = wfn_io.load(fh)
        if not (data["conserve_spin"] and data["conserve_number"]):
            raise NotImplementedError("number- or spin-broken wavefunctions are outside the "
                                      "B200 hot path")
        self._conserved, self._norb = dict(data["conserved"]), data["norb"]
        self._civec = {}
        for (nele, m_s), arr in data["sectors"].items():
            na, nb = alpha_beta_electrons(nele, m_s)
            sec = FqeData(na, nb, self._norb)
            sec.set_wfn(strategy='from_data', raw_data=arr)
            self._civec[(nele, m_s)] = sec

    # ---- individual n-body operators (wavefunction.py:1135-1328) ---------------------------
    def _operator_lists(self, alpha, beta):
        for oper in list(alpha) + list(beta):
            assert oper[0] < self._norb
        daga, undaga, dagb, undagb = sparse_hamiltonian.SparseHamiltonian.split(alpha, beta)
        if len(daga) + len(dagb) != len(undaga) + len(undagb):
            raise ValueError('Number non-conserving operators specified')
        if len(daga) != len(undaga) or len(dagb) != len(undagb):
            raise NotImplementedError("spin-changing individual operators need number-sector "
                                      "wavefunctions, which are outside the B200 hot path")
        return daga, undaga, dagb, undagb

    def _apply_individual_nbody(self, hamil: sparse_hamiltonian.SparseHamiltonian,
                                base: Optional['Wavefunction'] = None) -> 'Wavefunction':
        assert isinstance(hamil, sparse_hamiltonian.SparseHamiltonian)
        if hamil.nterms() > 1:
            raise ValueError('Indivisual n-body code is called with multiple terms')
        [(coeff, alpha, beta)] = hamil.terms()
        daga, undaga, dagb, undagb = self._operator_lists(alpha, beta)
        out = self.empty_copy() if base is None else base
        for key in self._civec.keys():
            out._civec[key].apply_individual_nbody_accumulate(coeff, self._civec[key], daga,
                                                              undaga, dagb, undagb)
        return out

    def _apply_few_nbody(self, hamil: sparse_hamiltonian.SparseHamiltonian) -> 'Wavefunction':
        out = None
        for oper in hamil.terms_hamiltonian():
            out = self._apply_individual_nbody(oper, base=out)
        if out is None:
            out = copy.deepcopy(self)
        if numpy.abs(hamil.e_0()) > 1.e-15:
            out.ax_plus_y(hamil.e_0(), self)
        return out

    def _evolve_individual_nbody(self, time: float, hamil: sparse_hamiltonian.SparseHamiltonian,
                                 inplace: bool = False) -> 'Wavefunction':
        """exp(-i t (T + T^+)) for one normal-ordered operator T given with (or as) its
        Hermitian conjugate (wavefunction.py:1192-1302)."""
        if not isinstance(hamil, sparse_hamiltonian.SparseHamiltonian):
            raise TypeError('Expected a Hamiltonian Object but received {}'.format(hamil))
        if hamil.nterms() > 2:
            raise ValueError('Individual n-body code is called with multiple terms')
        if hamil.nterms() == 2:
            [(coeff0, alpha0, beta0), (coeff1, alpha1, beta1)] = hamil.terms()
            check = all((a[0], a[1] ^ 1) in alpha1 for a in alpha0) and \
                all((b[0], b[1] ^ 1) in beta1 for b in beta0)
        else:
            [(coeff0, alpha0, beta0)] = hamil.terms()
            check = all((a[0], a[1] ^ 1) in alpha0 for a in alpha0) and \
                all((b[0], b[1] ^ 1) in beta0 for b in beta0)
        if not check:
            raise ValueError('Operators in _evolve_individual_nbody is not Hermitian')
        if hamil.nterms() == 1:
            coeff0 = coeff0 * 0.5
        daga, undaga, dagb, undagb = self._operator_lists(alpha0, beta0)
        if hamil.nterms() == 2:
            parity = (-1)**(len(alpha0) * len(beta0) + len(daga) * (len(daga) - 1) // 2 +
                            len(dagb) * (len(dagb) - 1) // 2 +
                            len(undaga) * (len(undaga) - 1) // 2 +
                            len(undagb) * (len(undagb) - 1) // 2)
            if not numpy.abs(coeff0 - numpy.conj(coeff1) * parity) < 1.0e-8:
                raise ValueError('Coefficients in _evolve_individual_nbody is not Hermitian')
        if daga == undaga and dagb == undagb:
            out = self if inplace else copy.deepcopy(self)
            for sector in out._civec.values():
                sector.evolve_inplace_individual_nbody_trivial(time, coeff0, daga, dagb)
        else:
            out = self.empty_copy(zero=False)
            for label, isec in self._civec.items():
                out._civec[label] = isec.evolve_individual_nbody_nontrivial(
                    time, coeff0, daga, undaga, dagb, undagb)
        return out

    def _evolve_diagonal(self, ithdiag: numpy.ndarray, inplace: bool = False) -> 'Wavefunction':
        """(wavefunction.py:1056-1077)"""
        wfn = self if inplace else copy.deepcopy(self)
        for sec in wfn._civec.values():
            sec.evolve_diagonal(ithdiag, inplace=True)
        return wfn

    def transform(self, rotation: numpy.ndarray, low: Optional[numpy.ndarray] = None,
                  upp: Optional[numpy.ndarray] = None):
        """Rotate the orbitals by the unitary ``rotation`` (norb x norb, the same for both
        spins, or a block-diagonal 2norb x 2norb).  Returns (permutation, L, U, self); the
        wavefunction is transformed IN PLACE like the reference's (wavefunction.py:813-959).

        rotation^H = P L U; the transformation operator factorises into 2*norb column
        operators, each applied by one pass of ``fqeb_apply_columns``."""
        norb = self._norb
        external = low is not None
        assert external == (upp is not None)
        if external:
            assert numpy.allclose(rotation, low @ upp)

        factors = rotation_factors

        if rotation.shape[0] == norb:
            perm, low, upp, output = factors(rotation, low, upp)
            for sec in self._civec.values():
                sec.apply_columns_recursive_inplace(output, output)
        elif rotation.shape[0] == 2 * norb:
            assert numpy.std(rotation[:norb, norb:]) + numpy.std(rotation[norb:, :norb]) < 1.0e-8
            la = None if low is None else low[:norb, :norb]
            ua = None if upp is None else upp[:norb, :norb]
            lb = None if low is None else low[norb:, norb:]
            ub = None if upp is None else upp[norb:, norb:]
            perm1, low1, upp1, output1 = factors(rotation[:norb, :norb], la, ua)
            perm2, low2, upp2, output2 = factors(rotation[norb:, norb:], lb, ub)
            for sec in self._civec.values():
                sec.apply_columns_recursive_inplace(output1, output2)
            if not external:
                perm, low, upp = (numpy.zeros_like(rotation), numpy.zeros_like(rotation),
                                  numpy.zeros_like(rotation))
                for full, blk1, blk2 in ((perm, perm1, perm2), (low, low1, low2),
                                         (upp, upp1, upp2)):
                    full[:norb, :norb] = blk1
                    full[norb:, norb:] = blk2
            else:
                perm = None
        else:
            raise ValueError("rotation must be norb x norb or 2norb x 2norb")
        return perm, low, upp, self

    def _evolve_diagonal_coulomb_inplace(self, diag: numpy.ndarray,
                                         vij: numpy.ndarray) -> 'Wavefunction':
        for sec in self._civec.values():
            sec.evolve_diagonal_coulomb(diag, vij, inplace=True)
        return self

    def expectationValue(self, ops, brawfn: Optional['Wavefunction'] = None):
        """<bra|H|self> via one sigma build and one dot product, or - for an operator string -
        an element (indices given as digits, openfermion convention) or a whole tensor (letters)
        of expectation values (wavefunction.py:1100-1133)."""
        if isinstance(ops, str):
            if any(ch.isdigit() for ch in ops):
                ops = sparse_hamiltonian.SparseHamiltonian(ops)
            else:
                return self.rdm(ops, brawfn=brawfn)
        if not isinstance(ops, hamiltonian.Hamiltonian):
            raise TypeError('Expected an Fqe Hamiltonian or Operator'
                            ' but recieved {}'.format(type(ops)))
        bra = brawfn if brawfn is not None else self
        return bra.vdot(self.apply(ops))

    def rdm(self, string: str, brawfn: Optional['Wavefunction'] = None):
        """Expectation values of the operator string (wavefunction.py:1331-1355): with digits
        (``'0^ 2'``, spin orbitals) one number <bra| op |self>; with letters (``'i^ j k l^'``) the
        spin-summed tensor over all spatial orbitals, assembled from the particle RDMs of the
        device path by Wick reordering (fqe_b200/wick.py)."""
        if any(ch.isdigit() for ch in string):
            result = self.apply(sparse_hamiltonian.SparseHamiltonian(string))
            return (self if brawfn is None else brawfn).vdot(result)
        tokens = string.split()
        if len(tokens) % 2 or not tokens:
            raise ValueError("an operator string needs an even number of operators")
        ncre = sum(1 for t in tokens if len(t) == 2 and t[0].islower() and t[1] == '^')
        nann = sum(1 for t in tokens if len(t) == 1 and t.islower())
        if ncre + nann != len(tokens):
            raise TypeError("Unsupported behavior for {}".format(string))
        if ncre != nann:
            raise ValueError("operator string does not conserve the particle number")
        from fqe_b200.wick import wick
        rank = len(tokens) // 2
        return wick(string, list(self._compute_rdm(rank, brawfn)), True)
