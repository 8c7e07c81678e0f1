"""CPU oracle for the FQE Hamiltonian-application path.  TEST INFRASTRUCTURE ONLY.

This module restates, in vectorised numpy, the *algorithm* of the reference's
pure-Python code path for the hot path named in BASELINE.json (string tables,
E_ij maps, dvec gather, two-electron contraction, coefficient scatter,
diagonal-Coulomb apply/evolve, Taylor / Chebyshev recurrences).  It is the
checker for the CUDA product in ``openfermion-fqe_b200`` and must never be
imported by product code: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

Parity status: PINNED.  ``tests/test_oracle_pinning.py`` checks this module
against (a) the reference's own golden vectors
(``/root/reference/tests/unittest_data/fqe_data/*.npy`` re-exported under
``tests/golden/`` by ``tests/golden/make_golden.py``), (b) the literal
known-answer tables in the reference's ``tests/fci_graph_test.py`` and (c) the
reference C library itself compiled from ``/root/reference`` into
``oracle/_ref`` (see ``oracle/ref_harness.py``).

All ``file:line`` citations are relative to ``/root/reference/src/fqe``.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Dict, Sequence, Tuple

import numpy as np
from scipy.special import comb, factorial, jv


# --------------------------------------------------------------------------
# a1: strings, Z matrix, addresses   (fci_graph.py:66-96, 301-333, 401-419)
# --------------------------------------------------------------------------
@lru_cache(maxsize=None)
def z_matrix(norb: int, nele: int) -> np.ndarray:
    """Knowles-Handy Z[k,l] addressing matrix (fci_graph.py:88-95,
    lib/fci_graph.c:27-54).  int32[nele, norb]."""
    z = np.zeros((nele, norb), dtype=np.int32)
    if z.size == 0:
        return z
    for k in range(1, nele):
        for ll in range(k, norb - nele + k + 1):
            acc = 0
            for m in range(norb - ll + 1, norb - k + 1):
                acc += comb(m, nele - k, exact=True) - comb(
                    m - 1, nele - k - 1, exact=True)
            z[k - 1, ll - 1] = acc
    for ll in range(nele, norb + 1):
        z[nele - 1, ll - 1] = ll - nele
    z.setflags(write=False)
    return z


def gosper_strings(nele: int, norb: int) -> np.ndarray:
    """All norb-bit strings with nele bits set in ascending integer order
    (bitstring.py lexicographic_bitstring_generator; lib/bitstring.c:29-46)."""
    n = comb(norb, nele, exact=True)
    out = np.zeros(n, dtype=np.uint64)
    if nele == 0:
        return out
    cur = (1 << nele) - 1
    lim = 1 << norb
    idx = 0
    while cur < lim:
        out[idx] = cur
        idx += 1
        low = cur & -cur
        ripple = cur + low
        cur = (((cur & ~ripple) // low) >> 1) | ripple
    assert idx == n
    return out


def occupations(strings: np.ndarray, norb: int) -> np.ndarray:
    """bool[L, norb] occupation table."""
    s = np.asarray(strings, dtype=np.uint64)
    return ((s[:, None] >> np.arange(norb, dtype=np.uint64)[None, :]) &
            np.uint64(1)).astype(bool)


def string_addresses(strings: np.ndarray, norb: int, nele: int) -> np.ndarray:
    """Z-matrix address of each string: sum_k Z[k, o_k] over its occupied
    orbitals o_0<o_1<...  (fci_graph.py:401-419, lib/fci_graph.c:123-134)."""
    strings = np.asarray(strings, dtype=np.uint64)
    if nele == 0:
        return np.zeros(strings.shape[0], dtype=np.int64)
    occ = occupations(strings, norb)
    # rank of each occupied orbital inside its string
    rank = np.cumsum(occ, axis=1) - 1
    z = z_matrix(norb, nele).astype(np.int64)
    orb = np.broadcast_to(np.arange(norb)[None, :], occ.shape)
    contrib = np.where(occ, z[np.clip(rank, 0, nele - 1), orb], 0)
    return contrib.sum(axis=1)


def build_strings(nele: int, norb: int) -> np.ndarray:
    """String table in Knowles-Handy address order (fci_graph.py:301-333)."""
    lex = gosper_strings(nele, norb)
    addr = string_addresses(lex, norb, nele)
    out = np.zeros_like(lex)
    out[addr] = lex
    return out


# --------------------------------------------------------------------------
# a2/a3: E_ij maps and de-excitation tables (fci_graph.py:204-237, 36-63)
# --------------------------------------------------------------------------
def _between_mask(i: int, j: int) -> int:
    lo, hi = (i, j) if i < j else (j, i)
    return ((1 << hi) - 1) & ~((1 << (lo + 1)) - 1)


def build_mapping(strings: np.ndarray, nele: int,
                  norb: int) -> Dict[Tuple[int, int], np.ndarray]:
    """(source, target, sign) triples of a^+_i a_j for every (i, j), sources in
    ascending string-index order (fci_graph.py:204-237; lib/fci_graph.c:76-121).
    sign = (-1)^(number of occupied orbitals strictly between i and j)
    (bitstring.py:206-221)."""
    strings = np.asarray(strings, dtype=np.uint64)
    nstr = strings.shape[0]
    sidx = np.arange(nstr, dtype=np.int64)
    out: Dict[Tuple[int, int], np.ndarray] = {}
    for i in range(norb):
        bi = np.uint64(1 << i)
        for j in range(norb):
            bj = np.uint64(1 << j)
            if i == j:
                sel = (strings & bj) != 0
                src = sidx[sel]
                trip = np.stack([src, src, np.ones_like(src)], axis=1)
            else:
                sel = ((strings & bj) != 0) & ((strings & bi) == 0)
                src = sidx[sel]
                s = strings[sel]
                tgt_str = (s | bi) & ~bj
                tgt = string_addresses(tgt_str, norb, nele)
                par = np.array(
                    [bin(int(x) & _between_mask(i, j)).count("1") for x in s],
                    dtype=np.int64)
                sign = 1 - 2 * (par & 1)
                trip = np.stack([src, tgt, sign], axis=1)
            out[(i, j)] = trip.astype(np.int32).reshape(-1, 3)
    return out


def map_to_deexc(mappings: Dict[Tuple[int, int], np.ndarray], nstates: int,
                 norb: int, nele: int) -> np.ndarray:
    """By-target table dexc[target, slot] = (source, i*norb+j, sign), slots
    filled in (i, j) row-major order (fci_graph.py:36-63)."""
    lk = nele * (norb - nele + 1)
    dexc = np.zeros((nstates, lk, 3), dtype=np.int32)
    fill = np.zeros(nstates, dtype=np.int64)
    for i in range(norb):
        for j in range(norb):
            m = mappings[(i, j)]
            if m.shape[0] == 0:
                continue
            tgt = m[:, 1]
            # a target appears at most once per (i, j)
            dexc[tgt, fill[tgt], 0] = m[:, 0]
            dexc[tgt, fill[tgt], 1] = i * norb + j
            dexc[tgt, fill[tgt], 2] = m[:, 2]
            fill[tgt] += 1
    return dexc


class OracleGraph:
    """Restatement of FciGraph.__init__ products (fci_graph.py:108-154)."""

    def __init__(self, nalpha: int, nbeta: int, norb: int):
        self.nalpha, self.nbeta, self.norb = nalpha, nbeta, norb
        self.lena = comb(norb, nalpha, exact=True)
        self.lenb = comb(norb, nbeta, exact=True)
        self.astr = build_strings(nalpha, norb)
        self.bstr = build_strings(nbeta, norb)
        self.alpha_map = build_mapping(self.astr, nalpha, norb)
        self.beta_map = build_mapping(self.bstr, nbeta, norb)
        self.dexca = map_to_deexc(self.alpha_map, self.lena, norb, nalpha)
        self.dexcb = map_to_deexc(self.beta_map, self.lenb, norb, nbeta)


@lru_cache(maxsize=32)
def graph(nalpha: int, nbeta: int, norb: int) -> OracleGraph:
    return OracleGraph(nalpha, nbeta, norb)


# --------------------------------------------------------------------------
# a7 / a8 / a9: dvec gather, contraction, coefficient scatter
# --------------------------------------------------------------------------
def dvec_spatial(g: OracleGraph, coeff: np.ndarray) -> np.ndarray:
    """D[i,j,t,:] += sign*C[s,:] (alpha maps), D[i,j,:,t] += sign*C[:,s]
    (beta maps)  (fqe_data.py:2209-2234)."""
    n = g.norb
    d = np.zeros((n, n, g.lena, g.lenb), dtype=np.complex128)
    for i in range(n):
        for j in range(n):
            ma = g.alpha_map[(i, j)]
            if ma.shape[0]:
                d[i, j, ma[:, 1], :] += coeff[ma[:, 0], :] * ma[:, 2][:, None]
            mb = g.beta_map[(i, j)]
            if mb.shape[0]:
                d[i, j, :, mb[:, 1]] += (coeff[:, mb[:, 0]] *
                                         mb[:, 2][None, :]).T
    return d


def coeff_from_dvec(g: OracleGraph, dvec: np.ndarray) -> np.ndarray:
    """out[s,:] += sign*D[i,j,t,:] with the maps of (j,i); columns for beta
    (fqe_data.py:2309-2334)."""
    n = g.norb
    out = np.zeros((g.lena, g.lenb), dtype=np.complex128)
    for i in range(n):
        for j in range(n):
            ma = g.alpha_map[(j, i)]
            if ma.shape[0]:
                np.add.at(out, (ma[:, 0], slice(None)),
                          dvec[i, j, ma[:, 1], :] * ma[:, 2][:, None])
            mb = g.beta_map[(j, i)]
            if mb.shape[0]:
                np.add.at(out, (slice(None), mb[:, 0]),
                          dvec[i, j][:, mb[:, 1]] * mb[:, 2][None, :])
    return out


def fold_restricted(h1: np.ndarray, h2: np.ndarray):
    """h2' = -moveaxis(h2,1,2); h1' = h1 - sum_k h2'[i,k,k,j]
    (fqe_data.py:647-652, 691-693)."""
    h2p = -np.moveaxis(np.asarray(h2, dtype=np.complex128), 1, 2)
    h1p = np.asarray(h1, dtype=np.complex128) - np.einsum("ikkj->ij", h2p)
    return h1p, np.ascontiguousarray(h2p)


def sigma_restricted(g: OracleGraph, coeff: np.ndarray, h1: np.ndarray,
                     h2: np.ndarray) -> np.ndarray:
    """FqeData.apply((h1,h2)) by the dvec route with complex dtype, which is
    valid for integrals of any symmetry class (fqe_data.py:653-657)."""
    h1p, h2p = fold_restricted(h1, h2)
    d = dvec_spatial(g, np.asarray(coeff, dtype=np.complex128))
    out = np.einsum("ij,ijkl->kl", h1p, d)
    n = g.norb
    e = (h2p.reshape(n * n, n * n) @ d.reshape(n * n, -1)).reshape(d.shape)
    out += coeff_from_dvec(g, e)
    return out


def fold_three_body(h1: np.ndarray, h2: np.ndarray, h3: np.ndarray):
    """Lower-rank pieces that normal-ordering the three-body term feeds into h1 and h2
    (fqe_data.py:1180-1191; the reference re-evaluates the 1+2-body apply inside its i loop,
    only the fully accumulated tensors of the last pass matter)."""
    n = h1.shape[0]
    nh1 = np.array(h1, dtype=np.complex128)
    nh2 = np.array(h2, dtype=np.complex128)
    for i in range(n):
        for j in range(n):
            for k in range(n):
                nh2[j, k, :, :] += (-h3[k, j, i, i, :, :] - h3[j, i, k, i, :, :] -
                                    h3[j, k, i, :, i, :])
            nh1[:, :] += h3[:, i, j, i, j, :]
    return nh1, nh2


def sigma_restricted_123(g: OracleGraph, coeff: np.ndarray, h1: np.ndarray, h2: np.ndarray,
                         h3: np.ndarray) -> np.ndarray:
    """FqeData.apply((h1,h2,h3)) = _apply_array_spatial123 (fqe_data.py:1166-1216):
    out = apply12(nh1, nh2) - scatter( sum_ij h3[:,q,i,:,s,j] . gather(gather(C)[i,j]) )."""
    n = g.norb
    coeff = np.asarray(coeff, dtype=np.complex128)
    nh1, nh2 = fold_three_body(h1, h2, np.asarray(h3, dtype=np.complex128))
    out = sigma_restricted(g, coeff, nh1, nh2)
    odvec = dvec_spatial(g, coeff)
    acc = np.zeros_like(odvec)
    for i in range(n):
        for j in range(n):
            tmp2 = dvec_spatial(g, odvec[i, j])
            acc += np.tensordot(h3[:, :, i, :, :, j], tmp2, axes=((1, 3), (0, 1)))
    return out - coeff_from_dvec(g, acc)


def sigma_one_body(g: OracleGraph, coeff: np.ndarray,
                   h1: np.ndarray) -> np.ndarray:
    """FqeData.apply((h1,)) (fqe_data.py:477-530)."""
    d = dvec_spatial(g, np.asarray(coeff, dtype=np.complex128))
    return np.einsum("ij,ijkl->kl", np.asarray(h1, dtype=np.complex128), d)


# --------------------------------------------------------------------------
# a10 / a11: diagonal Coulomb
# --------------------------------------------------------------------------
def dc_apply(g: OracleGraph, coeff: np.ndarray, diag: np.ndarray,
             array: np.ndarray) -> np.ndarray:
    """C[a,b] *= A[a]+B[b]+sum_{j in b} sum_{i in a}(v[i,j]+v[j,i])
    (fqe_data.py:293-330; lib/fqe_data.c:455-524)."""
    diag = np.asarray(diag, dtype=np.complex128)
    v = np.asarray(array, dtype=np.complex128)
    oa = occupations(g.astr, g.norb).astype(np.float64)
    ob = occupations(g.bstr, g.norb).astype(np.float64)
    a_d = oa @ diag + np.einsum("ai,ij,aj->a", oa, v, oa)
    b_d = ob @ diag + np.einsum("bi,ij,bj->b", ob, v, ob)
    cross = oa @ (v + v.T) @ ob.T
    return np.asarray(coeff, dtype=np.complex128) * (cross + a_d[:, None] +
                                                     b_d[None, :])


def dc_evolve(g: OracleGraph, coeff: np.ndarray, diag: np.ndarray,
              array: np.ndarray) -> np.ndarray:
    """C[a,b] *= Aexp[a]*Bexp[b]*(prod_{i in a, j in b} e^{v[i,j]})^2
    (fqe_data.py:363-402; lib/fqe_data.c:526-602).  Note the alpha-beta factor
    is squared, i.e. v is treated as symmetric."""
    de = np.exp(np.asarray(diag, dtype=np.complex128))
    ve = np.exp(np.asarray(array, dtype=np.complex128))

    def same_spin(strings):
        out = np.ones(strings.shape[0], dtype=np.complex128)
        occ = occupations(strings, g.norb)
        for s in range(strings.shape[0]):
            idx = np.nonzero(occ[s])[0]
            val = 1.0 + 0.0j
            for i in idx:
                val *= de[i]
                for j in idx:
                    val *= ve[i, j]
            out[s] = val
        return out, occ

    a_d, oa = same_spin(g.astr)
    b_d, ob = same_spin(g.bstr)
    out = np.array(coeff, dtype=np.complex128)
    for a in range(g.lena):
        rowprod = np.ones(g.norb, dtype=np.complex128)
        for i in np.nonzero(oa[a])[0]:
            rowprod *= ve[i, :]
        for b in range(g.lenb):
            x = 1.0 + 0.0j
            for j in np.nonzero(ob[b])[0]:
                x *= rowprod[j]
            out[a, b] *= x * x * a_d[a] * b_d[b]
    return out


def dc_tensors(h2e: np.ndarray):
    """DiagonalCoulomb.__init__: 2-D input -> (0, h2e); 4-D input ->
    diag[k]=h[k,k,k,k], vij[i,j]=-h[i,j,i,j] (hamiltonians/diagonal_coulomb.py:54-72)."""
    h2e = np.asarray(h2e)
    n = h2e.shape[0]
    if h2e.ndim == 2:
        return np.zeros(n, dtype=h2e.dtype), h2e
    diag = np.array([h2e[k, k, k, k] for k in range(n)], dtype=h2e.dtype)
    vij = np.zeros((n, n), dtype=h2e.dtype)
    for i in range(n):
        for j in range(n):
            vij[i, j] = -h2e[i, j, i, j]
    return diag, vij


# --------------------------------------------------------------------------
# a12-a15: Wavefunction-level semantics on one (n, sz) sector
# --------------------------------------------------------------------------
def wfn_apply_restricted(g, coeff, h1, h2, e0=0.0):
    """Wavefunction._apply_array: sigma, then += e_0*psi when |e_0|>1e-15
    (wavefunction.py:366-399)."""
    out = sigma_restricted(g, coeff, h1, h2)
    if abs(e0) > 1.0e-15:
        out = out + e0 * coeff
    return out


def wfn_apply_dc(g, coeff, diag, vij, e0=0.0):
    """Wavefunction._apply_diagonal_coulomb (wavefunction.py:421-440)."""
    out = dc_apply(g, coeff, diag, vij)
    if abs(e0) > 1.0e-15:
        out = out + e0 * coeff
    return out


def taylor(g, coeff, time, h1, h2, e0=0.0, accuracy=1.0e-15, expansion=30):
    """apply_generated_unitary(algo='taylor') (wavefunction.py:556-568, 604-605).
    Returns (state, nterms).  The tuple re-wrap inside the loop drops e_0
    (fqe_decorators.py:68-73), so e_0 only enters as the final phase."""
    ih1 = -1.0j * time * np.asarray(h1)
    ih2 = -1.0j * time * np.asarray(h2)
    evol = np.array(coeff, dtype=np.complex128)
    work = np.array(coeff, dtype=np.complex128)
    for order in range(1, expansion):
        work = sigma_restricted(g, work, ih1, ih2)
        c = 1.0 / factorial(order)
        evol += c * work
        if np.linalg.norm(work) * abs(c) < accuracy:
            break
    else:
        raise RuntimeError("maximum taylor expansion limit reached")
    if abs(e0 * time) > 1.0e-15:
        evol = evol * np.exp(-1.0j * time * e0)
    return evol, order


def chebyshev(g, coeff, time, h1, h2, spec_lim: Sequence[float], e0=0.0,
              accuracy=1.0e-15, expansion=30):
    """apply_generated_unitary(algo='chebyshev') (wavefunction.py:570-605).
    Here apply(hamil) carries e_0 (the Hamiltonian object itself is applied)."""
    wprime = 0.9875
    ascale = (spec_lim[1] - spec_lim[0]) / (2.0 * wprime)
    eshift = -(spec_lim[0] + ascale * wprime)

    def app(x):
        return wfn_apply_restricted(g, x, h1, h2, e0)

    base = np.array(coeff, dtype=np.complex128)
    evol = base * jv(0, ascale * time)
    minus = base.copy()
    current = app(minus)
    current = (current + eshift * minus) * (1.0 / ascale)
    evol = evol + 2.0 * jv(1, ascale * time) * (-1.0j) * current
    for order in range(2, expansion):
        minus = minus * (-1.0)
        minus = minus + (2.0 / ascale) * app(current)
        minus = minus + (2.0 * eshift / ascale) * current
        current, minus = minus, current
        c = 2.0 * jv(order, ascale * time) * (-1.0j)**order
        evol = evol + c * current
        if np.linalg.norm(current) * abs(c) < accuracy:
            break
    else:
        raise RuntimeError("maximum chebyshev expansion limit reached")
    evol = evol * np.exp(eshift * time * 1.0j)
    if abs(e0 * time) > 1.0e-15:
        evol = evol * np.exp(-1.0j * time * e0)
    return evol, order


def time_evolve_restricted(g, coeff, time, h1, h2, e0=0.0):
    """Wavefunction.time_evolve for a non-quadratic RestrictedHamiltonian:
    Taylor, then the e_0 phase a SECOND time (wavefunction.py:1042-1052;
    SURVEY F5)."""
    out, nterms = taylor(g, coeff, time, h1, h2, e0)
    if abs(e0) > 1.0e-15:
        out = out * np.exp(-1.0j * time * e0)
    return out, nterms


def time_evolve_dc(g, coeff, time, diag, vij, e0=0.0):
    """Wavefunction.time_evolve for DiagonalCoulomb: exact diagonal evolution
    with iht tensors, then the e_0 phase (wavefunction.py:1036-1040, 1051-1052)."""
    out = dc_evolve(g, coeff, -1.0j * time * np.asarray(diag),
                    -1.0j * time * np.asarray(vij))
    if abs(e0) > 1.0e-15:
        out = out * np.exp(-1.0j * time * e0)
    return out


# --------------------------------------------------------------------------
# independent check: dense matrix of H in the determinant basis
# --------------------------------------------------------------------------
# --------------------------------------------------------------------------
# 8f rank 1: orbital rotation by LU column operators, quadratic evolution
# (wavefunction.py:813-959, 1013-1034; fqe_data.py:1476-1535; lib/fqe_data.c:305-347)
# --------------------------------------------------------------------------
def apply_column(g: OracleGraph, coeff: np.ndarray, col: np.ndarray, icol: int,
                 spin: int) -> np.ndarray:
    """(1 + sum_i col[i] a^+_i a_icol) acting on one spin, out of place.  The reference does
    it in place: targets (icol empty) accumulate from sources (icol occupied), which are then
    scaled by 1 + col[icol] (fqe_data.py:1494-1504, lib/fqe_data.c:318-346)."""
    out = coeff.copy()
    maps = g.alpha_map if spin == 0 else g.beta_map
    for i in range(g.norb):
        m = maps[(i, icol)]
        if m.shape[0] == 0:
            continue
        if spin == 0:
            np.add.at(out, (m[:, 1], slice(None)), col[i] * coeff[m[:, 0], :] * m[:, 2][:, None])
        else:
            np.add.at(out, (slice(None), m[:, 1]), col[i] * coeff[:, m[:, 0]] * m[:, 2][None, :])
    return out


def apply_columns_recursive(g: OracleGraph, coeff: np.ndarray, mat1: np.ndarray,
                            mat2: np.ndarray) -> np.ndarray:
    """FqeData.apply_columns_recursive_inplace (fqe_data.py:1506-1535): all alpha columns in
    ascending order, then all beta columns."""
    c = np.asarray(coeff, dtype=np.complex128)
    for icol in range(g.norb):
        c = apply_column(g, c, mat1[:, icol], icol, 0)
    for icol in range(g.norb):
        c = apply_column(g, c, mat2[:, icol], icol, 1)
    return c


def lu_factors(rotation: np.ndarray):
    """ludecomp + transpose_matrix of Wavefunction.transform (wavefunction.py:838-887):
    P, L, U of rotation^H and their conjugate-transposed, diagonal-swapped counterparts."""
    from scipy import linalg
    perm, low, upp = linalg.lu(np.asarray(rotation).T.conj())
    lowt, uppt = low.copy(), upp.copy()
    n = low.shape[0]
    for irow in range(n):
        uppt[irow, irow + 1:] /= uppt[irow, irow]
        lowt[irow, irow], uppt[irow, irow] = uppt[irow, irow], lowt[irow, irow]
        for icol in range(irow):
            lowt[irow, icol] *= lowt[icol, icol]
    return perm, low, upp, uppt.T.conj(), lowt.T.conj()


def column_operator(low: np.ndarray, upp: np.ndarray) -> np.ndarray:
    """process_matrix (wavefunction.py:889-909): inv(upp) - strict_lower(low) - 1."""
    from scipy import linalg
    n = low.shape[0]
    out = linalg.solve_triangular(upp, np.identity(n))
    out = out - np.tril(low, -1) - np.identity(n)
    return out


def transform(g: OracleGraph, coeff: np.ndarray, rotation: np.ndarray, low=None, upp=None):
    """Wavefunction.transform for one number- and spin-conserving sector with a spatial
    rotation (wavefunction.py:918-928).  Returns (perm, low, upp, coeff')."""
    perm = None
    if low is None:
        perm, low, upp, lowt, uppt = lu_factors(rotation)
    else:
        lowt, uppt = low, upp
    mat = column_operator(lowt, uppt)
    return perm, low, upp, apply_columns_recursive(g, coeff, mat, mat)


def evolve_diagonal(g: OracleGraph, coeff: np.ndarray, array: np.ndarray) -> np.ndarray:
    """FqeData.evolve_diagonal (fqe_data.py:203-261): C[a,b] *= exp(sum_{i in a} array[i]) *
    exp(sum_{i in b} array[i])."""
    ea = np.exp(occupations(g.astr, g.norb) @ np.asarray(array, dtype=np.complex128))
    eb = np.exp(occupations(g.bstr, g.norb) @ np.asarray(array, dtype=np.complex128))
    return coeff * ea[:, None] * eb[None, :]


def apply_diagonal(g: OracleGraph, coeff: np.ndarray, array: np.ndarray) -> np.ndarray:
    """FqeData.apply_diagonal_inplace (fqe_data.py:170-201)."""
    da = occupations(g.astr, g.norb) @ np.asarray(array, dtype=np.complex128)
    db = occupations(g.bstr, g.norb) @ np.asarray(array, dtype=np.complex128)
    return coeff * (da[:, None] + db[None, :])


def time_evolve_quadratic(g: OracleGraph, coeff: np.ndarray, time: float, h1: np.ndarray,
                          e0=0.0) -> np.ndarray:
    """Quadratic branch of Wavefunction.time_evolve (wavefunction.py:1013-1034): rotate to the
    eigenbasis of h1, evolve the diagonal, rotate back with the same L, U factors."""
    h1 = np.asarray(h1, dtype=np.complex128)
    _, trans = np.linalg.eigh(h1)
    perm, low, upp, c = transform(g, coeff, trans)
    ci_trans = trans @ perm
    h1d = ci_trans.conj().T @ h1 @ ci_trans
    c = evolve_diagonal(g, c, -1.0j * time * h1d.diagonal())
    _, _, _, c = transform(g, c, ci_trans.T.conj(), low, upp)
    if abs(e0) > 1.0e-15:
        c = c * np.exp(-1.0j * time * e0)
    return c


# --------------------------------------------------------------------------
# 8f rank 2: individual n-body operators
# (fci_graph.py:511-570; fqe_data.py:1558-1653, 2385-2580; wavefunction.py:1135-1328)
# --------------------------------------------------------------------------
def make_mapping_each(strings: np.ndarray, norb: int, nele: int, dag, undag):
    """(source index, target index, sign) of prod a^+_dag prod a_undag on one spin; the
    rightmost operator acts first, sign = (-1)^(occupied orbitals above the one acted on)
    (fci_graph.py:546-569, lib/fci_graph.c:223-264)."""
    src, tgt, sgn = [], [], []
    for idx, s in enumerate(int(x) for x in strings):
        cur, parity, ok = s, 0, True
        for o in reversed(list(undag)):
            if not (cur >> o) & 1:
                ok = False
                break
            parity += bin(cur >> (o + 1)).count("1")
            cur &= ~(1 << o)
        if not ok:
            continue
        for o in reversed(list(dag)):
            if (cur >> o) & 1:
                ok = False
                break
            parity += bin(cur >> (o + 1)).count("1")
            cur |= 1 << o
        if not ok:
            continue
        src.append(idx)
        tgt.append(cur)
        sgn.append(1 - 2 * (parity & 1))
    tgt_idx = string_addresses(np.array(tgt, dtype=np.uint64), norb, nele) if tgt else []
    return (np.array(src, dtype=np.int64), np.array(tgt_idx, dtype=np.int64),
            np.array(sgn, dtype=np.int64))


def apply_individual_nbody(g: OracleGraph, coeff_in: np.ndarray, coeff: complex, daga, undaga,
                           dagb, undagb, out: np.ndarray = None) -> np.ndarray:
    """out[ta, tb] += coeff * pa * pb * in[sa, sb]  (fqe_data.py:1590-1667)."""
    if out is None:
        out = np.zeros_like(coeff_in, dtype=np.complex128)
    sa, ta, pa = make_mapping_each(g.astr, g.norb, g.nalpha, daga, undaga)
    sb, tb, pb = make_mapping_each(g.bstr, g.norb, g.nbeta, dagb, undagb)
    if sa.size and sb.size:
        out[np.ix_(ta, tb)] += coeff * (pa[:, None] * pb[None, :]) * coeff_in[np.ix_(sa, sb)]
    return out


def _occupied_mask(strings, occ, emp):
    pm = sum(1 << int(o) for o in set(occ))
    hm = sum(1 << int(o) for o in set(emp))
    return np.array([(int(s) & pm) == pm and (int(s) & hm) == 0 for s in strings], dtype=bool)


def sparse_scale(g: OracleGraph, coeff: np.ndarray, factor: complex, opa, oha, opb, ohb):
    """apply_cos_inplace / trivial evolution kernel: scale the determinants with opa, opb
    occupied and oha, ohb empty (fqe_data.py:2515-2580)."""
    ma, mb = _occupied_mask(g.astr, opa, oha), _occupied_mask(g.bstr, opb, ohb)
    out = coeff.copy()
    out[np.ix_(ma, mb)] *= factor
    return out


def evolve_individual_trivial(g, coeff, time, zc, opa, opb):
    """fqe_data.py:2385-2433"""
    n_a, n_b = len(opa), len(opb)
    zc = zc * (-1)**(n_a * (n_a - 1) // 2 + n_b * (n_b - 1) // 2)
    return sparse_scale(g, coeff, np.exp(-time * np.real(zc) * 2.j), opa, [], opb, [])


def evolve_individual_nontrivial(g, coeff, time, zc, daga, undaga, dagb, undagb):
    """exp(-i t (T + T^+)) with T^2 = 0 (fqe_data.py:2435-2513)."""
    def isolate(dag, undag, dagwork, undagwork, number):
        par = 0
        for cur in dag:
            if cur in undag:
                i1, i2 = dagwork.index(cur), undagwork.index(cur)
                par += len(dagwork) - (i1 + 1) + i2
                dagwork.remove(cur)
                undagwork.remove(cur)
                number.append(cur)
        return par

    dwa, uwa, dwb, uwb = list(daga), list(undaga), list(dagb), list(undagb)
    numa, numb = [], []
    parity = isolate(daga, undaga, dwa, uwa, numa) + isolate(dagb, undagb, dwb, uwb, numb)
    ncoeff = zc * (-1)**parity
    absol = abs(ncoeff)
    sinf = np.sin(time * absol) / absol
    out = sparse_scale(g, coeff, np.cos(time * absol), numa + dwa, uwa, numb + dwb, uwb)
    out = sparse_scale(g, out, np.cos(time * absol), numa + uwa, dwa, numb + uwb, dwb)
    phase = (-1)**((len(daga) + len(undaga)) * (len(dagb) + len(undagb)))
    apply_individual_nbody(g, coeff, np.conj(zc) * phase * (-1.0j) * sinf, undaga, daga, undagb,
                           dagb, out=out)
    apply_individual_nbody(g, coeff, zc * (-1.0j) * sinf, daga, undaga, dagb, undagb, out=out)
    return out


def split_operator(alpha, beta):
    daga = [o[0] for o in alpha if o[1] == 1]
    undaga = [o[0] for o in alpha if o[1] == 0]
    dagb = [o[0] for o in beta if o[1] == 1]
    undagb = [o[0] for o in beta if o[1] == 0]
    return daga, undaga, dagb, undagb


def sparse_apply(g, coeff, operators, e0=0.0):
    """Wavefunction._apply_few_nbody (wavefunction.py:1304-1328) for operators given in the
    SparseHamiltonian internal form (coeff, alpha ops, beta ops)."""
    out = np.zeros_like(coeff, dtype=np.complex128) if operators else coeff.copy()
    for zc, alpha, beta in operators:
        apply_individual_nbody(g, coeff, zc, *split_operator(alpha, beta), out=out)
    if abs(e0) > 1.0e-15:
        out = out + e0 * coeff
    return out


def ladder_sequence_apply(g: OracleGraph, coeff: np.ndarray, ops, zc: complex = 1.0):
    """INDEPENDENT brute force: apply zc * (product of spin-orbital ladder operators, in the
    order written, rightmost first) to the state.  ops = ((spin_orbital, 1|0), ...) with spin
    orbital 2k = (k, alpha), 2k+1 = (k, beta).  A determinant is
    [alpha creators, highest orbital leftmost][beta creators, highest leftmost]|0>, the
    convention behind make_mapping_each's count-bits-above parity and the alpha-then-beta
    string product (fci_graph.py:100-106).  Returns the projection back onto the sector."""
    state = {}
    for ia, sa in enumerate(int(x) for x in g.astr):
        for ib, sb in enumerate(int(x) for x in g.bstr):
            if coeff[ia, ib] != 0:
                state[(sa, sb)] = complex(coeff[ia, ib])
    for so, dagger in reversed(list(ops)):
        orb, beta = so // 2, so % 2
        new = {}
        for (sa, sb), amp in state.items():
            cur = sb if beta else sa
            occupied = (cur >> orb) & 1
            if occupied == dagger:
                continue
            par = bin(cur >> (orb + 1)).count("1")
            if beta:
                par += bin(sa).count("1")
            cur ^= 1 << orb
            key = (sa, cur) if beta else (cur, sb)
            new[key] = new.get(key, 0.0) + amp * (1 - 2 * (par & 1))
        state = new
    out = np.zeros((g.lena, g.lenb), dtype=np.complex128)
    aidx = {int(s): i for i, s in enumerate(g.astr)}
    bidx = {int(s): i for i, s in enumerate(g.bstr)}
    for (sa, sb), amp in state.items():
        if sa in aidx and sb in bidx:
            out[aidx[sa], bidx[sb]] += zc * amp
    return out


# --------------------------------------------------------------------------
# 8f rank 4: 1- and 2-particle (transition) RDMs  (fqe_data.py:1668-1838)
# --------------------------------------------------------------------------
def rdm12(g: OracleGraph, ket: np.ndarray, bra: np.ndarray = None):
    """FqeData._rdm12_halffilling, Python branch (fqe_data.py:1764-1780):
    rdm1[i,j] = <bra| a+_i a_j |ket>,  rdm2[i,j,k,l] = <bra| a+_i a+_j a_k a_l |ket>."""
    dvec = dvec_spatial(g, ket)
    dvec2 = dvec if bra is None else dvec_spatial(g, bra)
    out1 = np.transpose(np.tensordot(dvec2.conj(), ket))
    out2 = np.transpose(np.tensordot(dvec2.conj(), dvec, axes=((2, 3), (2, 3))),
                        axes=(1, 2, 0, 3)) * (-1.0)
    for i in range(g.norb):
        out2[:, i, i, :] += out1[:, :]
    return out1, out2


def dense_hamiltonian(g, h1, h2, e0=0.0):
    """Column-by-column H matrix, as tests/evolution_test.py:527-594 builds it."""
    dim = g.lena * g.lenb
    hmat = np.zeros((dim, dim), dtype=np.complex128)
    for k in range(dim):
        e = np.zeros(dim, dtype=np.complex128)
        e[k] = 1.0
        hmat[:, k] = wfn_apply_restricted(g, e.reshape(g.lena, g.lenb), h1, h2,
                                          e0).reshape(-1)
    return hmat


def rel_err(x, ref) -> float:
    x = np.asarray(x)
    ref = np.asarray(ref)
    den = np.linalg.norm(ref)
    return float(np.linalg.norm(x - ref) / (den if den > 0 else 1.0))
