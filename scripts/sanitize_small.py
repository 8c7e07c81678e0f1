"""Small end-to-end pass over every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import numpy as np
import torch
import fqe_b200 as fqe
from fqe_b200 import synth
from scipy.linalg import expm

for (n, sz, norb) in [(4, 0, 4), (5, 1, 6), (6, 0, 7), (8, 0, 8)]:
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): synth.state(la, lb, seed=norb)})
    for kind in ("real8", "herm"):
        h1, h2 = synth.integrals(norb, kind)
        ham = fqe.get_restricted_hamiltonian((h1, 0.1 * h2), e_0=0.1)
        for fusion in ("1", "0"):
            os.environ["FQEB_FUSION"] = fusion
            wfn.apply(ham)
        os.environ.pop("FQEB_FUSION")
        wfn.time_evolve(0.01, ham)
    wfn.time_evolve(0.2, fqe.get_restricted_hamiltonian((h1,)))
    a = np.random.default_rng(1).standard_normal((norb, norb))
    w2 = wfn.time_evolve(0.0, fqe.get_diagonal_hamiltonian(np.zeros(norb, dtype=complex)))
    w2.transform(expm(-0.2j * (a + a.T)))
    vij = synth.diagonal_coulomb_matrix(norb, 1)
    dc = fqe.get_diagonalcoulomb_hamiltonian(vij)
    wfn.apply(dc)
    wfn.time_evolve(0.1, dc)
    wfn.apply(fqe.get_diagonal_hamiltonian(np.arange(norb, dtype=complex)))
    sec = wfn.sector((n, sz))
    sec.apply_individual_nbody(0.3, [norb - 1], [0], [1], [2])
    sec.evolve_individual_nbody_nontrivial(0.2, 0.4 + 0.1j, [norb - 1], [0], [1], [2])
    sec.rdm12()
    sec.rdm12(w2.sector((n, sz)))                      # transition form of the Gram kernel
    wfn.rdm("i j^ k^ l")
    from fqe_b200.fqe_data import DenseOperator
    opd = DenseOperator(norb, h1, 0.1 * h2)
    for ozaki in ("1", "0"):                           # sliced and FP64 contraction, deferred scatter
        os.environ["FQEB_OZAKI"] = ozaki
        sig, pend = sec.apply_operator(opd, defer_last_scatter=True)
        half = sec.lena() // 2
        sec.finish_scatter(pend, 0, half, sig)
        sec.finish_scatter(pend, half, sec.lena(), sig)
    os.environ.pop("FQEB_OZAKI")
    wfn.norm(); wfn.vdot(wfn); wfn.scale(0.5); wfn.ax_plus_y(0.1, w2)
    if norb <= 6:
        h3 = 0.01 * np.ones((norb,) * 6, dtype=complex)
        wfn.apply((h1, 0.1 * h2, h3))
torch.cuda.synchronize()
print("sanitize_small: done")
