"""Host-side logic that needs no GPU: Hamiltonian containers, electron counting,
shard arithmetic, synthetic inputs."""
import os
import numpy as np
import pytest

from fqe_b200 import synth
from fqe_b200.distributed import shard_plan, split_even
from fqe_b200.hamiltonians import diagonal_coulomb, restricted_hamiltonian
from fqe_b200.wavefunction import alpha_beta_electrons, build_hamiltonian
from oracle import fqe_oracle as O


def test_alpha_beta_electrons():
    assert alpha_beta_electrons(4, 0) == (2, 2)
    assert alpha_beta_electrons(5, 1) == (3, 2)
    assert alpha_beta_electrons(3, -3) == (0, 3)
    for bad in [(-1, 0), (2, 4), (3, 0)]:
        with pytest.raises(ValueError):
            alpha_beta_electrons(*bad)


def test_restricted_hamiltonian_container():
    h1, h2 = synth.integrals(4, "herm")
    ham = restricted_hamiltonian.RestrictedHamiltonian((h1, h2), e_0=1.5)
    assert ham.dim() == 4 and ham.rank() == 4 and not ham.quadratic()
    assert ham.e_0() == 1.5 and ham.conserve_number()
    assert not ham.diagonal() and not ham.diagonal_coulomb()
    t = ham.tensors()
    assert t[0] is h1 and t[1] is h2
    i1, i2 = ham.iht(0.3)
    assert np.array_equal(i1, -0.3j * h1) and np.array_equal(i2, -0.3j * h2)
    assert ham == restricted_hamiltonian.RestrictedHamiltonian((h1.copy(), h2.copy()), e_0=1.5)
    assert ham != restricted_hamiltonian.RestrictedHamiltonian((h1, h2), e_0=0.0)
    quad = restricted_hamiltonian.RestrictedHamiltonian((h1,))
    assert quad.quadratic() and quad.rank() == 2
    w, v = np.linalg.eigh(h1)
    assert np.allclose(quad.transform(quad.calc_diag_transform()), np.diag(w))
    with pytest.raises(TypeError):
        restricted_hamiltonian.RestrictedHamiltonian(([1, 2],))
    with pytest.raises(ValueError):
        restricted_hamiltonian.RestrictedHamiltonian((np.zeros((2, 2, 2)),))


def test_diagonal_coulomb_container_matches_oracle():
    rng = np.random.default_rng(3)
    v = rng.standard_normal((5, 5))
    dc = diagonal_coulomb.DiagonalCoulomb(v, e_0=-0.5)
    assert dc.diagonal_coulomb() and dc.rank() == 4 and dc.dim() == 5
    assert np.array_equal(dc._tensor[1], np.zeros(5)) and dc._tensor[2] is v
    h4 = rng.standard_normal((5,) * 4)
    dc4 = diagonal_coulomb.DiagonalCoulomb(h4)
    diag, vij = O.dc_tensors(h4)
    assert np.array_equal(dc4._tensor[1], diag) and np.array_equal(dc4._tensor[2], vij)
    d, a = dc.iht(0.1)
    assert np.array_equal(a, -0.1j * v) and np.array_equal(d, np.zeros(5))
    with pytest.raises(ValueError):
        diagonal_coulomb.DiagonalCoulomb(np.zeros((2, 2, 2)))


def test_build_hamiltonian_dispatch():
    h1, h2 = synth.integrals(3, "real8")
    ham = build_hamiltonian((h1, h2), norb=3)
    assert isinstance(ham, restricted_hamiltonian.RestrictedHamiltonian) and ham.e_0() == 0
    assert build_hamiltonian(ham) is ham
    with pytest.raises(NotImplementedError):
        build_hamiltonian((np.zeros((6, 6)),), norb=3)
    with pytest.raises(TypeError):
        build_hamiltonian("a+ a", norb=3)
    with pytest.raises(TypeError):
        build_hamiltonian(([1.0],), norb=3)


def test_split_even_and_shard_plan():
    for total, world, align in [(12870, 8, 1), (136, 8, 16), (256, 8, 16), (36, 4, 2), (7, 3, 1),
                                (5, 8, 1)]:
        parts = split_even(total, world, align)
        assert parts[0][0] == 0 and parts[-1][1] == total
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and a <= b
        for lo, hi in parts[:-1]:
            assert hi % align == 0 or hi == total
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(s for s in sizes) <= align + align
    rows, pairs = shard_plan("det", 3, 8, 12870, 136)
    assert pairs == (0, 136) and rows == split_even(12870, 8)[3]
    rows, pairs = shard_plan("pair", 1, 8, 12870, 256)
    assert rows == (0, 12870) and pairs == (32, 64)
    rows, pairs = shard_plan("pair", 7, 8, 12870, 136)
    assert pairs[1] == 136 and pairs[0] % 2 == 0
    with pytest.raises(ValueError):
        shard_plan("rows", 0, 2, 10, 4)


def test_synthetic_integrals_symmetry_classes():
    h1, h2 = synth.integrals(5, "real8")
    assert h1.dtype == np.complex128 and np.all(h1.imag == 0) and np.array_equal(h1, h1.T)
    h2p = -np.moveaxis(h2, 1, 2)
    assert np.array_equal(h2p, h2p.transpose(1, 0, 2, 3))   # pair symmetric -> compressible
    assert np.array_equal(h2p, h2p.transpose(0, 1, 3, 2))
    assert np.array_equal(h2p, h2p.transpose(2, 3, 0, 1))
    g = O.graph(2, 2, 5)
    hm = O.dense_hamiltonian(g, h1, h2)
    assert np.allclose(hm, hm.conj().T)
    h1, h2 = synth.integrals(4, "herm")
    g = O.graph(2, 1, 4)
    hm = O.dense_hamiltonian(g, h1, h2)
    assert np.allclose(hm, hm.conj().T) and np.abs(hm.imag).max() > 1e-3
    c = synth.state(6, 4, seed=1)
    assert abs(np.linalg.norm(c) - 1) < 1e-14 and np.array_equal(c, synth.state(6, 4, seed=1))


def test_diagonal_hamiltonian_container():
    """Diagonal mirrors hamiltonians/diagonal_hamiltonian.py:26-121"""
    import fqe_b200
    from fqe_b200.hamiltonians.diagonal_hamiltonian import Diagonal
    d = np.array([0.5, -1.0, 2.0], dtype=np.complex128)
    h = fqe_b200.get_diagonal_hamiltonian(d, e_0=0.25)
    assert isinstance(h, Diagonal)
    assert h.dim() == 3 and h.rank() == 2 and h.quadratic() and h.diagonal()
    assert not h.diagonal_coulomb() and h.conserve_number() and h.e_0() == 0.25
    assert np.array_equal(h.diag_values(), d)
    it = h.iht(0.1)
    assert np.allclose(it.diag_values(), -0.1j * d) and np.array_equal(h.diag_values(), d)
    assert h == Diagonal(d.copy(), e_0=0.25) and not (h == Diagonal(d + 1, e_0=0.25))
    with pytest.raises(ValueError):
        Diagonal(np.zeros((2, 2)))


def test_quadratic_transform_helpers():
    """RestrictedHamiltonian.calc_diag_transform / transform diagonalise h1
    (restricted_hamiltonian.py:136-155)"""
    import fqe_b200
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5))
    h1 = a + a.conj().T
    ham = fqe_b200.get_restricted_hamiltonian((h1,))
    assert ham.quadratic() and not ham.diagonal()
    u = ham.calc_diag_transform()
    hd = ham.transform(u)
    assert np.allclose(hd, np.diag(np.diag(hd)), atol=1e-12)
    assert np.allclose(np.sort(np.diag(hd).real), np.linalg.eigvalsh(h1))


def test_sparse_hamiltonian_normal_ordering_against_brute_force():
    """SparseHamiltonian's own normal ordering + alpha/beta split (openfermion is not available
    here): the operator it encodes must act like the ladder-operator product as written"""
    from fqe_b200.hamiltonians.sparse_hamiltonian import SparseHamiltonian, normal_ordered
    assert normal_ordered({((1, 0), (2, 1)): 1.0}) == {((2, 1), (1, 0)): -1.0}
    assert normal_ordered({((1, 0), (1, 1)): 1.0}) == {(): 1.0, ((1, 1), (1, 0)): -1.0}
    assert normal_ordered({((3, 1), (3, 1)): 1.0}) == {}
    g = O.graph(3, 2, 5)
    rng = np.random.default_rng(9)
    c = rng.standard_normal((g.lena, g.lenb)) + 1j * rng.standard_normal((g.lena, g.lenb))
    cases = ["3^ 0 2^ 1", "0^ 2", "5 1^ 1 5^", "4^ 2^ 0 6", "1^ 3^ 7 5", "2 2^",
             "6^ 1^ 3^ 3 7 0", "9^ 8 8^ 9", "0 4^ 3^ 7"]
    for text in cases:
        ham = SparseHamiltonian(text)
        seq = tuple((int(t.rstrip('^')), 1 if t.endswith('^') else 0) for t in text.split())
        ref = O.ladder_sequence_apply(g, c, seq)
        ops = ham.terms()
        spin_ok = all(sum(1 for o in a if o[1]) * 2 == len(a) and
                      sum(1 for o in b if o[1]) * 2 == len(b) for _, a, b in ops)
        if not spin_ok:
            continue   # spin-changing products leave the sector; only the constructor is checked
        out = O.sparse_apply(g, c, ops, ham.e_0())
        assert np.abs(out - ref).max() < 1e-13, text
    # a multi-term mapping with a constant: e_0 picks it up
    ham = SparseHamiltonian({((0, 1), (2, 0)): 0.5, ((2, 1), (0, 0)): 0.5, (): 1.5}, e_0=0.25)
    assert ham.e_0() == 1.75 and ham.nterms() == 2 and ham.is_individual() and ham.rank() == 2
    assert not SparseHamiltonian({((0, 1), (2, 0)): 0.5, ((2, 1), (0, 0)): 0.5,
                                  ((4, 1), (4, 0)): 1.0}).is_individual()
    it = ham.iht(0.5)
    assert it.terms()[0][0] == -0.25j and ham.terms()[0][0] == 0.5


def test_rotation_factors_match_oracle_and_reconstruct():
    """host part of Wavefunction.transform: LU factors of rot^H and the column-operator matrix"""
    from scipy.linalg import expm
    from fqe_b200.wavefunction import rotation_factors
    rng = np.random.default_rng(4)
    for n in (2, 5, 8):
        a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        rot = expm(-0.4j * (a + a.conj().T))
        perm, low, upp, mat = rotation_factors(rot)
        assert np.allclose(perm @ low @ upp, rot.conj().T, atol=1e-13)
        operm, olow, oupp, lowt, uppt = O.lu_factors(rot)
        assert np.array_equal(perm, operm) and np.allclose(low, olow) and np.allclose(upp, oupp)
        assert np.allclose(mat, O.column_operator(lowt, uppt), atol=1e-13)
        # the way back: external factors are taken as they are
        back = (rot @ perm).conj().T
        assert np.allclose(back, low @ upp, atol=1e-13)
        _, l2, u2, mat2 = rotation_factors(back, low, upp)
        assert l2 is low and u2 is upp
        assert np.allclose(mat2, O.column_operator(low, upp), atol=1e-13)


def test_wick_reordering_matches_reference(golden_dir):
    """fqe_b200.wick against the reference's Wavefunction.rdm outputs (tests/golden/ref_wick.npz,
    generated by make_golden.py --only wick): every string is rebuilt from the two particle RDMs
    the same file holds ('i^ j' and 'i^ j^ k l')."""
    from fqe_b200.wick import wick
    g = np.load(os.path.join(golden_dir, "ref_wick.npz"))
    strings = [str(s) for s in g["strings"]]
    for tag in ("wa", "wb", "wc"):
        for pre in ("s", "t"):
            data = [g[f"{tag}_{pre}0"], g[f"{tag}_{pre}3"]]
            for k, st in enumerate(strings):
                ref = g[f"{tag}_{pre}{k}"]
                got = wick(st, data[:len(st.split()) // 2])
                assert got.shape == ref.shape
                assert np.abs(got - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), (tag, st)


def test_wick_rejects_bad_strings():
    from fqe_b200.wick import wick
    r1 = np.zeros((3, 3), dtype=np.complex128)
    r2 = np.zeros((3,) * 4, dtype=np.complex128)
    with pytest.raises(ValueError):
        wick("i^ j^ k^ l", [r1, r2])     # two creators in spin slot 0
    with pytest.raises(ValueError):
        wick("i^ jj", [r1])
    with pytest.raises(ValueError):
        wick("i^ j k", [r1, r2])
    with pytest.raises(ValueError):
        wick("i^ j^ k l", [r1])          # rank-2 string needs rdm2
