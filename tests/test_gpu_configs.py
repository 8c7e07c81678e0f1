"""BASELINE.json configurations as parity cases.

config 0: H-ring 12 orbitals / 12 electrons, Sz=0 (reference profiling/profile_H_ring.py):
          integrals from profiling/Hring_12.hdf5, Hartree-Fock state, expectationValue,
          time_evolve(0.1) - compared with values recorded from the reference itself
          (tests/golden/ref_hring12.npz, written by tests/golden/make_golden.py).
config 3: a Trotter step of a double-factorised Hamiltonian = diagonal-Coulomb evolution
          composed with an orbital rotation exp(-i dt K) (quadratic time_evolve), against
          exact exponentials built with the oracle.
Tolerance: 1e-10 relative (north_star)."""
import os

import numpy as np
import pytest

from oracle import fqe_oracle as O

pytestmark = pytest.mark.gpu


def test_config0_hring12(golden_dir):
    import fqe_b200 as fqe
    g = np.load(os.path.join(golden_dir, "ref_hring12.npz"))
    nele, sz, norbs = [int(x) for x in g["meta"]]
    ham = fqe.get_restricted_hamiltonian((g["h1"], g["h2"]), e_0=float(g["e_0"][0]))
    wf = fqe.Wavefunction([[nele, sz, norbs]])
    wf.set_wfn(strategy="hartree-fock")
    wf.normalize()
    e_init = wf.expectationValue(ham)
    assert abs(e_init - complex(g["e_init"][0])) < 1e-10 * abs(g["e_init"][0])
    assert abs(e_init.real - float(g["hf_energy"][0])) < 1e-9   # the file's own HF energy

    shape = wf.get_coeff((nele, sz)).shape
    rng = np.random.default_rng(int(g["probe_seed"][0]))
    probe = rng.standard_normal((8,) + shape) + 1j * rng.standard_normal((8,) + shape)
    idx = g["sample_idx"]

    def check(state, tag):
        nrm = float(g[f"{tag}_norm"][0])
        assert abs(np.linalg.norm(state) - nrm) < 1e-10 * nrm
        assert np.linalg.norm(state.reshape(-1)[idx] - g[f"{tag}_samples"]) < 1e-10 * nrm
        proj = np.einsum("kab,ab->k", probe.conj(), state)
        assert np.linalg.norm(proj - g[f"{tag}_probe"]) < 1e-10 * np.linalg.norm(g[f"{tag}_probe"])

    check(wf.apply(ham).get_coeff((nele, sz)), "sigma")
    evolved = wf.time_evolve(float(g["t"][0]), ham)
    check(evolved.get_coeff((nele, sz)), "evolved")
    e_final = evolved.expectationValue(ham)
    assert abs(e_final - complex(g["e_final"][0])) < 1e-10 * abs(g["e_final"][0])
    assert abs(e_final.real - e_init.real) < 1e-10 * abs(e_init.real)   # energy conserved


def test_config3_trotter_step_small():
    """exp(-i dt V_dc) exp(-i dt K) on a random state, norb=6: both factors exact"""
    import fqe_b200 as fqe
    from fqe_b200 import synth
    norb, n, sz = 6, 6, 0
    g = O.graph(3, 3, norb)
    c0 = synth.state(g.lena, g.lenb, seed=8)
    rng = np.random.default_rng(12)
    k = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
    k = 0.5 * (k + k.conj().T)
    vij = synth.diagonal_coulomb_matrix(norb, 5)
    dt = 0.05
    wf = fqe.Wavefunction([[n, sz, norb]])
    wf.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    rot = wf.time_evolve(dt, fqe.get_restricted_hamiltonian((k,)))
    out = rot.time_evolve(dt, fqe.get_diagonalcoulomb_hamiltonian(vij)).get_coeff((n, sz))
    # oracle: dense one-body matrix exponential, then the diagonal factor
    dim = c0.size
    hm = np.zeros((dim, dim), dtype=np.complex128)
    for col in range(dim):
        e = np.zeros(dim, dtype=np.complex128)
        e[col] = 1.0
        hm[:, col] = O.sigma_one_body(g, e.reshape(c0.shape), k).reshape(-1)
    w, v = np.linalg.eigh(hm)
    ref = ((v * np.exp(-1j * dt * w)) @ (v.conj().T @ c0.reshape(-1))).reshape(c0.shape)
    ref = O.time_evolve_dc(g, ref, dt, np.zeros(norb), vij)
    assert O.rel_err(out, ref) < 1e-10
    assert abs(np.linalg.norm(out) - 1.0) < 1e-12


@pytest.mark.parametrize("tag", ["t3a", "t3b", "t3c"])
def test_config5_dense_three_body_matches_reference_api(golden_dir, tag):
    """wfn.apply((h1, h2, h3)) against outputs of the reference's own API; t3c is the tensor
    of profiling/profile_3_body.py (h1 = h2 = 0)"""
    import fqe_b200 as fqe
    api = np.load(os.path.join(golden_dir, "ref_api.npz"))
    n, sz, norb = [int(x) for x in api[f"{tag}_meta"]]
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): api[f"{tag}_c0"]})
    out = wfn.apply((api[f"{tag}_h1"], api[f"{tag}_h2"], api[f"{tag}_h3"])).get_coeff((n, sz))
    assert O.rel_err(out, api[f"{tag}_sigma"]) < 1e-11
    out = wfn.sector((n, sz)).apply((api[f"{tag}_h1"], api[f"{tag}_h2"], api[f"{tag}_h3"]))
    assert O.rel_err(out.to_numpy(), api[f"{tag}_sigma"]) < 1e-11


def test_config5_three_body_vs_oracle_norb7():
    import fqe_b200 as fqe
    from fqe_b200 import synth
    norb, n, sz = 7, 7, 1
    g = O.graph(4, 3, norb)
    rng = np.random.default_rng(55)
    h1 = rng.standard_normal((norb,) * 2)
    h2 = 0.1 * rng.standard_normal((norb,) * 4)
    h3 = 0.02 * (rng.standard_normal((norb,) * 6) + 1j * rng.standard_normal((norb,) * 6))
    c = synth.state(g.lena, g.lenb, seed=9)
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c})
    out = wfn.apply((h1, h2, h3)).get_coeff((n, sz))
    assert O.rel_err(out, O.sigma_restricted_123(g, c, h1, h2, h3)) < 1e-11


def test_three_body_time_evolve_small():
    """Taylor evolution under a Hermitian 1+2+3-body operator against the exact exponential"""
    import fqe_b200 as fqe
    from fqe_b200 import synth
    norb, n, sz = 4, 4, 0
    g = O.graph(2, 2, norb)
    rng = np.random.default_rng(77)
    h1, h2 = synth.integrals(norb, "herm")
    w = rng.standard_normal((norb,) * 6) + 1j * rng.standard_normal((norb,) * 6)
    h3 = 0.02 * (w + w.conj().transpose(5, 4, 3, 2, 1, 0))   # a+a+a+ a a a hermitian
    c = synth.state(g.lena, g.lenb, seed=2)
    dim = c.size
    hm = np.zeros((dim, dim), dtype=np.complex128)
    for col in range(dim):
        e = np.zeros(dim, dtype=np.complex128)
        e[col] = 1.0
        hm[:, col] = O.sigma_restricted_123(g, e.reshape(c.shape), h1, h2, h3).reshape(-1)
    assert np.allclose(hm, hm.conj().T, atol=1e-12)
    wv, vv = np.linalg.eigh(hm)
    t = 0.05
    ref = ((vv * np.exp(-1j * t * wv)) @ (vv.conj().T @ c.reshape(-1))).reshape(c.shape)
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c})
    out = wfn.time_evolve(t, fqe.get_restricted_hamiltonian((h1, h2, h3)))
    assert O.rel_err(out.get_coeff((n, sz)), ref) < 1e-10
