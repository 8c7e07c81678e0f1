"""SURVEY 8f rank 2 on the GPU: individual n-body apply and exact evolution (FqeData level and
through Wavefunction + SparseHamiltonian), against outputs of the reference recorded in
tests/golden/ref_nbody.npz, the oracle, and an independent brute-force ladder-operator walk.
Tolerance 1e-12 relative 2-norm (target 1e-10); applies are signed copies and must be exact."""
import ast
import os

import numpy as np
import pytest

from oracle import fqe_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12
TAGS = ["na", "nb", "nc", "nd"]


@pytest.fixture(scope="module")
def nbody(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_nbody.npz"))


def _setup(z, tag):
    import fqe_b200
    n, sz, norb = [int(x) for x in z[f"{tag}_meta"]]
    wfn = fqe_b200.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): z[f"{tag}_c0"]})
    return fqe_b200, wfn, (n, sz), norb


@pytest.mark.parametrize("tag", TAGS)
def test_fqedata_individual_nbody(nbody, tag):
    fqe, wfn, key, norb = _setup(nbody, tag)
    ops = [ast.literal_eval(str(o)) for o in nbody["ops"]]
    zc, time = complex(nbody["coeff"][0]), float(nbody["time"][0])
    sec = wfn.sector(key)
    for k, (da, ua, db, ub) in enumerate(ops):
        if f"{tag}_apply{k}" not in nbody:
            continue
        out = sec.apply_individual_nbody(zc, da, ua, db, ub)
        assert np.abs(out.to_numpy() - nbody[f"{tag}_apply{k}"]).max() < 1e-15
        assert np.array_equal(sec.to_numpy(), nbody[f"{tag}_c0"])
        if f"{tag}_trivial{k}" in nbody:
            tmp = fqe.Wavefunction([[key[0], key[1], norb]])
            tmp.set_wfn(strategy="from_data", raw_data={key: nbody[f"{tag}_c0"]})
            tmp.sector(key).evolve_inplace_individual_nbody_trivial(time, zc, da, db)
            assert O.rel_err(tmp.get_coeff(key), nbody[f"{tag}_trivial{k}"]) < TOL
        else:
            ev = sec.evolve_individual_nbody_nontrivial(time, zc, da, ua, db, ub)
            assert O.rel_err(ev.to_numpy(), nbody[f"{tag}_evolve{k}"]) < TOL
            assert abs(np.linalg.norm(ev.to_numpy()) - 1.0) < 1e-12   # unitary


@pytest.mark.parametrize("tag", TAGS)
def test_wavefunction_sparse_hamiltonian(nbody, tag):
    if f"{tag}_w_pair_apply" not in nbody:
        pytest.skip("sector too small for the recorded operators")
    from fqe_b200.hamiltonians.sparse_hamiltonian import SparseHamiltonian
    fqe, wfn, key, norb = _setup(nbody, tag)
    zc, time = complex(nbody["coeff"][0]), float(nbody["time"][0])
    t_op = (zc, [(2, 1), (0, 0)], [(1, 1), (3, 0)])
    t_dag = (np.conj(zc), [(0, 1), (2, 0)], [(3, 1), (1, 0)])
    pair = SparseHamiltonian.from_operators([t_op, t_dag], e_0=0.25)
    assert pair.is_individual()
    assert O.rel_err(wfn.apply(pair).get_coeff(key), nbody[f"{tag}_w_pair_apply"]) < TOL
    assert O.rel_err(wfn.time_evolve(time, pair).get_coeff(key), nbody[f"{tag}_w_pair_evolve"]) < TOL
    num = SparseHamiltonian.from_operators([(0.7, [(1, 1), (1, 0)], [(0, 1), (0, 0)])], e_0=-0.5)
    assert O.rel_err(wfn.apply(num).get_coeff(key), nbody[f"{tag}_w_num_apply"]) < TOL
    assert O.rel_err(wfn.time_evolve(time, num).get_coeff(key), nbody[f"{tag}_w_num_evolve"]) < TOL
    three = SparseHamiltonian.from_operators([t_op, t_dag, (0.4, [(1, 1), (1, 0)], [])], e_0=0.05)
    assert not three.is_individual()
    assert O.rel_err(wfn.apply(three).get_coeff(key), nbody[f"{tag}_w_three_apply"]) < TOL
    assert O.rel_err(wfn.time_evolve(0.05, three).get_coeff(key),
                     nbody[f"{tag}_w_three_evolve"]) < TOL
    assert O.rel_err(wfn.get_coeff(key), nbody[f"{tag}_c0"]) == 0.0


def test_sparse_hamiltonian_from_fermion_operator_terms():
    """front end: terms mapping / string in the FermionOperator convention -> GPU apply and
    evolve agree with the brute-force walk of the operator product as written and with the
    exact exponential of T + T^+"""
    import fqe_b200
    n, sz, norb = 5, 1, 5
    g = O.graph(3, 2, norb)
    rng = np.random.default_rng(77)
    c0 = rng.standard_normal((g.lena, g.lenb)) + 1j * rng.standard_normal((g.lena, g.lenb))
    c0 /= np.linalg.norm(c0)
    wfn = fqe_b200.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    for text in ["3^ 0 2^ 1", "0^ 2", "4^ 2^ 0 6", "1^ 3^ 7 5", "6^ 1^ 3^ 3 7 0", "9^ 8 8^ 9"]:
        ham = fqe_b200.get_sparse_hamiltonian(text)
        seq = tuple((int(t.rstrip('^')), 1 if t.endswith('^') else 0) for t in text.split())
        ref = O.ladder_sequence_apply(g, c0, seq)
        assert np.abs(wfn.apply(ham).get_coeff((n, sz)) - ref).max() < 1e-14, text
    # T + T^+ with T = 0.6 e^{0.3i} a+_{2a} a_{0a} a+_{1b} a_{3b} (spin orbitals 4^ 0 3^ 7)
    zc = 0.6 * np.exp(0.3j)
    fwd = ((4, 1), (0, 0), (3, 1), (7, 0))
    bwd = ((7, 1), (3, 0), (0, 1), (4, 0))
    ham = fqe_b200.get_sparse_hamiltonian({fwd: zc, bwd: np.conj(zc)})
    assert ham.is_individual()
    dim = g.lena * g.lenb
    hm = np.zeros((dim, dim), dtype=np.complex128)
    for k in range(dim):
        e = np.zeros(dim, dtype=np.complex128)
        e[k] = 1.0
        e = e.reshape(g.lena, g.lenb)
        hm[:, k] = (O.ladder_sequence_apply(g, e, fwd, zc) +
                    O.ladder_sequence_apply(g, e, bwd, np.conj(zc))).reshape(-1)
    assert np.abs(hm - hm.conj().T).max() < 1e-14
    w, v = np.linalg.eigh(hm)
    t = 0.83
    exact = (v * np.exp(-1j * t * w)) @ (v.conj().T @ c0.reshape(-1))
    out = wfn.time_evolve(t, ham).get_coeff((n, sz))
    assert O.rel_err(out.reshape(-1), exact) < 1e-12
    # the Taylor route gives the same state
    tay = wfn.apply_generated_unitary(t, "taylor", ham).get_coeff((n, sz))
    assert O.rel_err(tay.reshape(-1), exact) < 1e-12


def test_nbody_error_paths():
    import fqe_b200
    from fqe_b200.hamiltonians.sparse_hamiltonian import SparseHamiltonian
    wfn = fqe_b200.Wavefunction([[4, 0, 4]])
    wfn.set_wfn(strategy="ones")
    sec = wfn.sector((4, 0))
    with pytest.raises(NotImplementedError):
        sec.apply_individual_nbody(1.0, [1], [], [], [0])       # spin flip
    with pytest.raises(ValueError):
        sec.apply_individual_nbody(1.0, [7], [0], [], [])       # orbital out of range
    with pytest.raises(ValueError):
        sec.apply_individual_nbody_accumulate(1.0, sec, [1], [0], [], [])   # aliased
    with pytest.raises(ValueError):
        wfn.time_evolve(0.1, SparseHamiltonian.from_operators(
            [(1.0, [(1, 1), (0, 0)], []), (0.5, [(2, 1), (1, 0)], [])]))    # not Hermitian


def test_nbody_properties_at_scale():
    """norb=12 half filling (853776 determinants): a two-body excitation generator evolves
    unitarily, and evolving forward then backward in time is the identity"""
    import fqe_b200
    from fqe_b200 import synth
    n, sz, norb = 12, 0, 12
    c0 = synth.state(924, 924, seed=31)
    wfn = fqe_b200.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    zc = 0.8 - 0.3j
    fwd = ((2 * 9, 1), (2 * 2, 0), (2 * 7 + 1, 1), (2 * 4 + 1, 0))
    bwd = ((2 * 4 + 1, 1), (2 * 7 + 1, 0), (2 * 2, 1), (2 * 9, 0))
    ham = fqe_b200.get_sparse_hamiltonian({fwd: zc, bwd: np.conj(zc)})
    ev = wfn.time_evolve(0.6, ham)
    assert abs(ev.norm() - 1.0) < 1e-12
    assert O.rel_err(ev.get_coeff((n, sz)), c0) > 1e-3
    back = ev.time_evolve(-0.6, ham)
    assert O.rel_err(back.get_coeff((n, sz)), c0) < 1e-12
