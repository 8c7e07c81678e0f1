"""SURVEY 8f rank 4 on the GPU: 1- and 2-particle (transition) RDMs and save / read, against
the reference's FqeData.rdm1 / rdm12 outputs (tests/golden/ref_rdm.npz) and the oracle."""
import os

import numpy as np
import pytest

from oracle import fqe_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def rdm(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_rdm.npz"))


@pytest.mark.parametrize("tag", ["ra", "rb", "rc", "rd", "re"])
def test_rdm12_matches_reference(rdm, tag):
    import fqe_b200
    n, sz, norb = [int(x) for x in rdm[f"{tag}_meta"]]
    ket = fqe_b200.Wavefunction([[n, sz, norb]])
    ket.set_wfn(strategy="from_data", raw_data={(n, sz): rdm[f"{tag}_ket"]})
    bra = fqe_b200.Wavefunction([[n, sz, norb]])
    bra.set_wfn(strategy="from_data", raw_data={(n, sz): rdm[f"{tag}_bra"]})
    sec, bsec = ket.sector((n, sz)), bra.sector((n, sz))
    (r1,) = sec.rdm1()
    assert O.rel_err(r1, rdm[f"{tag}_rdm1"]) < TOL
    (t1,) = sec.rdm1(bsec)
    assert O.rel_err(t1, rdm[f"{tag}_trdm1"]) < TOL
    r1, r2 = sec.rdm12()
    assert O.rel_err(r1, rdm[f"{tag}_rdm12_1"]) < TOL and O.rel_err(r2, rdm[f"{tag}_rdm12_2"]) < TOL
    t1, t2 = sec.rdm12(bsec)
    assert O.rel_err(t1, rdm[f"{tag}_trdm12_1"]) < TOL
    assert O.rel_err(t2, rdm[f"{tag}_trdm12_2"]) < TOL
    w1, w2 = ket._compute_rdm(2, bra)
    assert np.array_equal(w1, t1) and np.array_equal(w2, t2)
    # energy through the RDMs equals the expectation value through sigma
    from fqe_b200 import synth
    h1, h2 = synth.integrals(norb, "herm", seed=3)
    ham = fqe_b200.get_restricted_hamiltonian((h1, h2))
    e_sigma = ket.expectationValue(ham)
    e_rdm = np.einsum("ij,ij", h1, r1) + np.einsum("ijkl,ijkl", h2, r2)
    assert abs(e_sigma - e_rdm) < 1e-11 * max(1.0, abs(e_sigma))


def test_rdm_blocked_at_scale_and_save_read(tmp_path, monkeypatch):
    """norb=10 half filling through the blocked path (7 alpha-row blocks): trace and
    symmetry properties; save / read round trip"""
    import fqe_b200
    from fqe_b200 import synth
    from fqe_b200.fqe_data import FqeData
    monkeypatch.setattr(FqeData, "_rdm_block_bytes", 16 * 100 * 252 * 40)
    n, sz, norb = 10, 0, 10
    c0 = synth.state(252, 252, seed=8)
    wfn = fqe_b200.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    r1, r2 = wfn.sector((n, sz)).rdm12()
    assert abs(np.trace(r1) - n) < 1e-11
    assert np.abs(r1 - r1.conj().T).max() < 1e-13
    # sum_ij <a+_i(s) a+_j(t) a_i(s) a_j(t)> = -N(N-1): a_k carries the spin of a+_i
    assert abs(np.einsum("ijij", r2) + n * (n - 1)) < 1e-10
    g = O.graph(5, 5, norb)
    o1, o2 = O.rdm12(g, c0)
    assert O.rel_err(r1, o1) < TOL and O.rel_err(r2, o2) < TOL
    wfn.save("state.pkl", path=str(tmp_path))
    back = fqe_b200.Wavefunction([[n, sz, norb]])
    back.read("state.pkl", path=str(tmp_path))
    assert np.array_equal(back.get_coeff((n, sz)), c0)
