"""SURVEY 8f rank 1 on the GPU: Wavefunction.transform (LU column rotations), the quadratic
branch of time_evolve, Diagonal apply / evolve and FqeData.evolve_diagonal, against outputs of
the reference's public API (tests/golden/ref_transform.npz, made by make_golden.py) and against
the oracle.  Tolerance 1e-12 relative 2-norm (target 1e-10)."""
import copy
import os

import numpy as np
import pytest

from oracle import fqe_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12
TAGS = ["ta", "tb", "tc", "td", "te", "tf"]


@pytest.fixture(scope="module")
def rot(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_transform.npz"))


def _setup(z, tag):
    import fqe_b200
    n, sz, norb = [int(x) for x in z[f"{tag}_meta"]]
    wfn = fqe_b200.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): z[f"{tag}_c0"]})
    return fqe_b200, wfn, (n, sz), norb, float(z[f"{tag}_t"][0]), complex(z[f"{tag}_e0"][0])


@pytest.mark.parametrize("tag", TAGS)
def test_transform_matches_reference_api(rot, tag):
    fqe, wfn, key, norb, t, e0 = _setup(rot, tag)
    perm, low, upp, out = wfn.transform(rot[f"{tag}_rot"])
    assert out is wfn  # in place, like the reference
    assert np.array_equal(perm, rot[f"{tag}_perm"])
    assert np.allclose(low, rot[f"{tag}_low"], atol=1e-14, rtol=0)
    assert np.allclose(upp, rot[f"{tag}_upp"], atol=1e-14, rtol=0)
    assert O.rel_err(out.get_coeff(key), rot[f"{tag}_transformed"]) < TOL


@pytest.mark.parametrize("tag", TAGS)
def test_quadratic_evolve_and_apply(rot, tag):
    fqe, wfn, key, norb, t, e0 = _setup(rot, tag)
    ham = fqe.get_restricted_hamiltonian((rot[f"{tag}_h1"],), e_0=e0)
    out = wfn.time_evolve(t, ham)
    assert O.rel_err(out.get_coeff(key), rot[f"{tag}_quad_evolve"]) < TOL
    assert O.rel_err(wfn.get_coeff(key), rot[f"{tag}_c0"]) == 0.0   # out of place
    assert O.rel_err(wfn.apply(ham).get_coeff(key), rot[f"{tag}_quad_apply"]) < TOL
    # in-place evolution returns the same object
    w2 = copy.deepcopy(wfn)
    out2 = w2.time_evolve(t, ham, inplace=True)
    assert out2 is w2
    assert O.rel_err(out2.get_coeff(key), rot[f"{tag}_quad_evolve"]) < TOL
    # and agrees with the Taylor propagator of the same operator (half the time step so that
    # the series converges within its 30-term cap; the e_0 phase enters both routes once)
    half = wfn.time_evolve(0.1 * t, ham)
    bare = fqe.get_restricted_hamiltonian((rot[f"{tag}_h1"],), e_0=0.0)
    tay = wfn.apply_generated_unitary(0.1 * t, "taylor", bare)
    tay.scale(np.exp(-1j * 0.1 * t * e0))
    assert O.rel_err(tay.get_coeff(key), half.get_coeff(key)) < 1e-11


@pytest.mark.parametrize("tag", TAGS)
def test_diagonal_hamiltonian(rot, tag):
    fqe, wfn, key, norb, t, e0 = _setup(rot, tag)
    diag = rot[f"{tag}_diag"]
    dham = fqe.get_diagonal_hamiltonian(diag.astype(np.complex128), e_0=e0)
    assert O.rel_err(wfn.apply(dham).get_coeff(key), rot[f"{tag}_diag_apply"]) < TOL
    assert O.rel_err(wfn.time_evolve(t, dham).get_coeff(key), rot[f"{tag}_diag_evolve"]) < TOL
    sec = wfn.sector(key)
    ev = sec.evolve_diagonal(-1j * t * diag)
    assert O.rel_err(ev.cpu().numpy(), rot[f"{tag}_evolve_diagonal"]) < TOL
    assert O.rel_err(sec.to_numpy(), rot[f"{tag}_c0"]) == 0.0
    # separate alpha / beta arrays (2*norb)
    g = O.graph(sec.nalpha(), sec.nbeta(), norb)
    both = np.concatenate([diag, 0.5 - diag])
    ev = sec.evolve_diagonal(-1j * t * both)
    ea = np.exp(O.occupations(g.astr, norb) @ (-1j * t * diag))
    eb = np.exp(O.occupations(g.bstr, norb) @ (-1j * t * (0.5 - diag)))
    assert O.rel_err(ev.cpu().numpy(), rot[f"{tag}_c0"] * ea[:, None] * eb[None, :]) < TOL
    with pytest.raises(ValueError):
        sec.evolve_diagonal(np.zeros(norb + 1))


@pytest.mark.parametrize("cfg", [(1, 1, 2), (2, 3, 6), (3, 0, 5), (0, 2, 4), (4, 4, 4), (4, 1, 7),
                                 (5, 5, 10)])
def test_columns_vs_oracle(cfg):
    """general (non-unitary) column operators, different for alpha and beta"""
    from fqe_b200 import synth
    from fqe_b200.fqe_data import FqeData
    na, nb, norb = cfg
    g = O.graph(na, nb, norb)
    rng = np.random.default_rng(100 * norb + na)
    c = synth.state(g.lena, g.lenb, seed=7 + norb)
    m1 = 0.3 * (rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb)))
    m2 = 0.3 * (rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb)))
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    d.apply_columns_recursive_inplace(m1, m2)
    assert O.rel_err(d.to_numpy(), O.apply_columns_recursive(g, c, m1, m2)) < TOL


def test_spin_block_rotation_and_external_factors():
    """2norb x 2norb block-diagonal rotation (different alpha / beta unitaries), and the
    externally supplied L, U factors used on the way back in time_evolve"""
    import fqe_b200
    from scipy.linalg import expm, lu
    n, sz, norb = 5, 1, 6
    rng = np.random.default_rng(42)

    def unitary():
        a = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
        return expm(-0.6j * (a + a.conj().T))

    ua, ub = unitary(), unitary()
    big = np.zeros((2 * norb, 2 * norb), dtype=np.complex128)
    big[:norb, :norb], big[norb:, norb:] = ua, ub
    g = O.graph(3, 2, norb)
    c0 = rng.standard_normal((g.lena, g.lenb)) + 1j * rng.standard_normal((g.lena, g.lenb))
    wfn = fqe_b200.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    perm, low, upp, out = wfn.transform(big)
    ma = O.column_operator(*O.lu_factors(ua)[3:])
    mb = O.column_operator(*O.lu_factors(ub)[3:])
    assert O.rel_err(out.get_coeff((n, sz)), O.apply_columns_recursive(g, c0, ma, mb)) < TOL
    assert np.allclose(perm[:norb, :norb], lu(ua.T.conj())[0])
    # external factors: rotation == low @ upp is required
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    p, l, u = lu(ua.T.conj())
    back = (ua @ p).T.conj()
    _, _, _, out = wfn.transform(back, l, u)
    ref = O.transform(g, c0, back, l, u)[3]
    assert O.rel_err(out.get_coeff((n, sz)), ref) < TOL


def test_rotation_round_trip_at_scale():
    """norb=12 half filling: transform(U) followed by the rotation back with (U P)^H and the
    same L, U factors (the sequence time_evolve runs) is the identity and preserves the norm
    (size-independent property; 853776 determinants)"""
    import fqe_b200
    from fqe_b200 import synth
    from scipy.linalg import expm
    n, sz, norb = 12, 0, 12
    rng = np.random.default_rng(12)
    a = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
    u = expm(-0.3j * (a + a.conj().T))
    c0 = synth.state(924, 924, seed=5)
    wfn = fqe_b200.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    perm, low, upp, _ = wfn.transform(u)
    assert abs(wfn.norm() - 1.0) < 1e-12
    assert O.rel_err(wfn.get_coeff((n, sz)), c0) > 1e-2
    wfn.transform((u @ perm).T.conj(), low, upp)
    assert O.rel_err(wfn.get_coeff((n, sz)), c0) < 1e-11
