"""Operator strings -> tensors of expectation values, by Wick reordering of the particle RDMs.

Host-side front end of ``Wavefunction.rdm("i^ j k l^")`` (reference: src/fqe/wick.py:37-197 and
its C filler lib/wick.c).  The heavy part - the particle RDMs themselves - comes from the device
(FqeData.rdm12 -> csrc/rdm.cu); what is left is index algebra on [norb]^(2 rank) tensors, done
here with numpy.einsum.

Conventions reproduced from the reference (they define the meaning of a string):

* operators are single letters, ``x^`` is a creator; a string of 2r operators yields a tensor
  whose axes follow the operators in the order written;
* spin-free strings: the operator at position p carries the spin slot ``p % r``, i.e. the indices
  are read as 1 2 ... r / 1 2 ... r; two creators (or two annihilators) in the same slot make the
  string non-number-conserving and are rejected;
* a contraction between an annihilator and the creator to its right contributes a Kronecker
  delta; inside one spin slot the spin sum gives a factor 2, across slots the creator's slot is
  renamed to the annihilator's in the operators that are left;
* the particle RDMs are ordered creators-then-annihilators with ascending slots on both sides;
  bringing a term to that order costs a sign per transposition.
"""
from typing import List, NamedTuple, Sequence, Tuple

import numpy


class _Op(NamedTuple):
    label: str
    creator: bool
    slot: int


class _Term(NamedTuple):
    factor: float
    ops: Tuple[_Op, ...]                 # normal-ordered product that is left
    deltas: Tuple[Tuple[str, str], ...]  # pairs of labels tied by a Kronecker delta


def _parse(target: str, spinfree: bool) -> List[_Op]:
    tokens = target.split()
    if len(tokens) % 2:
        raise ValueError("an operator string needs an even number of operators")
    rank = len(tokens) // 2
    ops: List[_Op] = []
    for pos, tok in enumerate(tokens):
        creator = tok.endswith("^")
        label = tok[:-1] if creator else tok
        if len(label) != 1:
            raise ValueError('unrecognized input in wick')
        if any(o.label == label for o in ops):
            raise ValueError(f"index {label!r} appears twice in {target!r}")
        slot = pos % rank if spinfree else 0
        if spinfree and any(o.slot == slot and o.creator == creator for o in ops):
            raise ValueError('non-number conserving input to Wick')
        ops.append(_Op(label, creator, slot))
    return ops


def _normal_order(ops: Tuple[_Op, ...], factor: float, deltas: Tuple[Tuple[str, str], ...],
                  spinfree: bool, done: List[_Term]) -> None:
    """a_p a+_q = delta_pq - a+_q a_p, applied to the leftmost (annihilator, creator) neighbours
    until every creator stands left of every annihilator."""
    for p in range(len(ops) - 1):
        left, right = ops[p], ops[p + 1]
        if left.creator or not right.creator:
            continue
        swapped = ops[:p] + (right, left) + ops[p + 2:]
        _normal_order(swapped, -factor, deltas, spinfree, done)
        rest = ops[:p] + ops[p + 2:]
        weight = factor
        if spinfree:
            if left.slot == right.slot:
                weight *= 2.0
            else:
                rest = tuple(o._replace(slot=left.slot) if o.slot == right.slot else o
                             for o in rest)
        _normal_order(rest, weight, deltas + ((left.label, right.label),), spinfree, done)
        return
    done.append(_Term(factor, ops, deltas))


def _inversions(values: Sequence[int]) -> int:
    return sum(1 for i in range(len(values)) for j in range(i + 1, len(values))
               if values[i] > values[j])


def _slot_sorted(term: _Term) -> _Term:
    """creators and annihilators each sorted by slot (stable), with the permutation's sign"""
    half = len(term.ops) // 2
    cre, ann = term.ops[:half], term.ops[half:]
    swaps = _inversions([o.slot for o in cre]) + _inversions([o.slot for o in ann])
    ordered = tuple(sorted(cre, key=lambda o: o.slot)) + tuple(sorted(ann, key=lambda o: o.slot))
    return _Term(-term.factor if swaps % 2 else term.factor, ordered, term.deltas)


def wick(target: str, data: Sequence[numpy.ndarray], spinfree: bool = True) -> numpy.ndarray:
    """Tensor of <target> from the particle RDMs ``data = [rdm1, rdm2, ...]`` (as returned by
    ``Wavefunction._compute_rdm``); same result as the reference's ``fqe.wick.wick``."""
    ops = _parse(target, spinfree)
    rank = len(ops) // 2
    if rank < 1 or len(data) < rank:
        raise ValueError("wick needs the particle RDMs up to the rank of the string")
    terms: List[_Term] = []
    _normal_order(tuple(ops), 1.0, (), spinfree, terms)
    if spinfree:
        terms = [_slot_sorted(t) for t in terms]

    norb = data[rank - 1].shape[0]
    axis = {o.label: "abcdefgh"[pos] for pos, o in enumerate(ops)}
    out_sub = "".join(axis[o.label] for o in ops)
    out = numpy.zeros_like(data[rank - 1])
    eye = numpy.eye(norb, dtype=out.dtype)
    for term in terms:
        subs, operands = [], []
        if term.ops:
            subs.append("".join(axis[o.label] for o in term.ops))
            operands.append(data[len(term.ops) // 2 - 1])
        for x, y in term.deltas:
            subs.append(axis[x] + axis[y])
            operands.append(eye)
        out = out + term.factor * numpy.einsum(",".join(subs) + "->" + out_sub, *operands)
    return out
