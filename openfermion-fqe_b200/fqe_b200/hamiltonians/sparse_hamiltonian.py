"""SparseHamiltonian: a few individual products of ladder operators.

API-compatible with /root/reference/src/fqe/hamiltonians/sparse_hamiltonian.py:29-200 for
number- and spin-conserving terms.  The reference takes an ``openfermion.FermionOperator``
and normal-orders it with ``openfermion.transforms.normal_ordered``; openfermion is not a
dependency here, so the constructor accepts

* any object with a ``terms`` mapping in the FermionOperator convention
  ``{((spin_orbital, 1|0), ...): coefficient}`` (an actual FermionOperator works unchanged),
* such a mapping itself, or
* a single-term string ``"3^ 0 2^ 1"`` (coefficient 1, as ``FermionOperator(str, 1.0)``),

and normal-orders it with the same rule (creators to the left, each group by descending
spin-orbital index, a sign per transposition, a contraction for equal indices).  Spin orbital
2k is spatial orbital k with alpha spin, 2k+1 with beta spin (openfermion ``up_index`` /
``down_index``).  The split into alpha / beta blocks follows
hamiltonians/hamiltonian_utils.py:159-202 (``gather_nbody_spin_sectors``).
"""
import copy
from typing import Dict, List, Tuple

from fqe_b200.hamiltonians import hamiltonian

Term = Tuple[Tuple[int, int], ...]
Operator = Tuple[complex, List[Tuple[int, int]], List[Tuple[int, int]]]


def _parse_term_string(text: str) -> Term:
    ops = []
    for tok in text.split():
        if tok.endswith('^'):
            ops.append((int(tok[:-1]), 1))
        else:
            ops.append((int(tok), 0))
    return tuple(ops)


def _as_terms(operators) -> Dict[Term, complex]:
    if isinstance(operators, str):
        return {_parse_term_string(operators): 1.0}
    terms = getattr(operators, 'terms', operators)
    if not hasattr(terms, 'items'):
        raise TypeError("SparseHamiltonian expects a FermionOperator-like object, a terms "
                        "mapping or a term string")
    return {tuple((int(i), int(a)) for i, a in key): val for key, val in terms.items()}


def _normal_ordered_term(term: Term, coeff: complex, out: Dict[Term, complex]) -> None:
    """Creators left of annihilators, each group in descending index order (the rule of
    openfermion's normal_ordered_ladder_term); results accumulate into ``out``."""
    term = list(term)
    for i in range(1, len(term)):
        for j in range(i, 0, -1):
            right, left = term[j], term[j - 1]
            if right[1] and not left[1]:
                # a_p a+_q = delta_pq - a+_q a_p
                term[j - 1], term[j] = right, left
                coeff = -coeff
                if right[0] == left[0]:
                    _normal_ordered_term(tuple(term[:j - 1] + term[j + 1:]), -coeff, out)
            elif right[1] == left[1]:
                if right[0] == left[0]:
                    return  # a a = a+ a+ = 0
                if right[0] > left[0]:
                    term[j - 1], term[j] = right, left
                    coeff = -coeff
    key = tuple(term)
    out[key] = out.get(key, 0.0) + coeff


def normal_ordered(terms: Dict[Term, complex]) -> Dict[Term, complex]:
    out: Dict[Term, complex] = {}
    for term, coeff in terms.items():
        _normal_ordered_term(term, coeff, out)
    return {k: v for k, v in out.items() if v != 0.0}


def gather_nbody_spin_sectors(term: Term, coeff: complex):
    """(coeff, phase, alpha ops, beta ops) of one normal-ordered term: stable sort alpha (even
    spin orbitals) before beta, then creators and annihilators of each spin by descending
    index, counting transpositions (hamiltonian_utils.py:159-202, util.py:53-78, 199-225)."""
    ops = [list(o) for o in term]
    nswaps = 0
    n = len(ops)
    for i in range(n):            # bubble sort on the spin (parity of the index)
        swapped = False
        for j in range(n - i - 1):
            if ops[j][0] % 2 > ops[j + 1][0] % 2:
                ops[j], ops[j + 1] = ops[j + 1], ops[j]
                nswaps += 1
                swapped = True
        if not swapped:
            break

    def descending(block):
        cnt = 0
        m = len(block)
        for i in range(m):
            swapped = False
            for j in range(m - i - 1):
                if block[j][0] < block[j + 1][0]:
                    block[j], block[j + 1] = block[j + 1], block[j]
                    cnt += 1
                    swapped = True
            if not swapped:
                break
        return cnt

    nda = sum(1 for o in ops if o[0] % 2 == 0 and o[1] == 1)
    nalpha = sum(1 for o in ops if o[0] % 2 == 0)
    ndb = sum(1 for o in ops if o[0] % 2 == 1 and o[1] == 1)
    ablock, bblock = ops[:nalpha], ops[nalpha:]
    parts = [ablock[:nda], ablock[nda:], bblock[:ndb], bblock[ndb:]]
    for part in parts:
        nswaps += descending(part)
    alpha = [tuple(o) for o in parts[0] + parts[1]]
    beta = [tuple(o) for o in parts[2] + parts[3]]
    return coeff, (-1)**nswaps, alpha, beta


class SparseHamiltonian(hamiltonian.Hamiltonian):

    def __init__(self, operators, conserve_spin: bool = True, e_0: complex = 0.0 + 0.0j) -> None:
        terms = normal_ordered(_as_terms(operators))
        work = terms.pop((), None)
        if work is not None:
            e_0 += work
        super().__init__(e_0=e_0)
        self._operators: List[Operator] = []
        self._conserve_spin = conserve_spin
        self._rank = 0
        for prod in terms:
            self._rank = max(self._rank, len(prod))
        for prod, val in terms.items():
            coeff, phase, alpha_block, beta_block = gather_nbody_spin_sectors(prod, val)
            alpha_out = [(a[0] // 2, a[1]) for a in alpha_block]
            beta_out = [(b[0] // 2, b[1]) for b in beta_block]
            self._operators.append((coeff * phase, alpha_out, beta_out))

    @classmethod
    def from_operators(cls, operators: List[Operator], conserve_spin: bool = True,
                       e_0: complex = 0.0 + 0.0j) -> 'SparseHamiltonian':
        """Build directly from (coeff, [(orbital, 1|0) alpha ops], [(orbital, 1|0) beta ops])
        triples, the internal form ``terms()`` returns (sparse_hamiltonian.py:170-178)."""
        out = cls({}, conserve_spin=conserve_spin, e_0=e_0)
        out._operators = [(c, list(a), list(b)) for c, a, b in operators]
        out._rank = max([len(a) + len(b) for _, a, b in out._operators] + [0])
        return out

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, SparseHamiltonian):
            return NotImplemented
        return self.e_0() == other.e_0() and self._conserve_spin == other._conserve_spin \
            and self._operators == other._operators

    def dim(self):
        raise NotImplementedError

    def rank(self) -> int:
        return self._rank

    def nterms(self) -> int:
        return len(self._operators)

    @staticmethod
    def split(alpha, beta):
        """(daga, undaga, dagb, undagb) orbital lists of one operator, in product order"""
        daga = [o[0] for o in alpha if o[1] == 1]
        undaga = [o[0] for o in alpha if o[1] == 0]
        dagb = [o[0] for o in beta if o[1] == 1]
        undagb = [o[0] for o in beta if o[1] == 0]
        return daga, undaga, dagb, undagb

    def is_individual(self) -> bool:
        """True for one operator plus (at most) its Hermitian conjugate"""
        nterm = 0
        for (_, alpha, beta) in self._operators:
            daga, undaga, dagb, undagb = self.split(alpha, beta)
            nterm += 2 if (daga == undaga and dagb == undagb) else 1
        return nterm < 3

    def iht(self, time: float) -> 'SparseHamiltonian':
        out = copy.deepcopy(self)
        out._operators = [(-coeff * 1.0j * time, alpha, beta)
                          for coeff, alpha, beta in out._operators]
        return out

    def terms(self) -> List[Operator]:
        return self._operators

    def terms_hamiltonian(self) -> List['SparseHamiltonian']:
        out = []
        for current in self._operators:
            tmp = copy.deepcopy(self)
            tmp._operators = [current]
            out.append(tmp)
        return out
