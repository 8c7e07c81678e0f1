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


def test_rdm_strings_match_reference(golden_dir):
    """Wavefunction.rdm / expectationValue with operator strings: every spin-free rank-1 / rank-2
    ordering against the reference's own outputs (tests/golden/ref_wick.npz), RDMs from the
    device path"""
    import fqe_b200
    g = np.load(os.path.join(golden_dir, "ref_wick.npz"))
    strings = [str(s) for s in g["strings"]]
    for tag in ("wa", "wb", "wc"):
        n, sz, norb = [int(x) for x in g[f"{tag}_meta"]]
        ket = fqe_b200.Wavefunction([[n, sz, norb]])
        ket.set_wfn(strategy="from_data", raw_data={(n, sz): g[f"{tag}_ket"]})
        bra = fqe_b200.Wavefunction([[n, sz, norb]])
        bra.set_wfn(strategy="from_data", raw_data={(n, sz): g[f"{tag}_bra"]})
        for k, st in enumerate(strings):
            assert O.rel_err(ket.rdm(st), g[f"{tag}_s{k}"]) < TOL, (tag, st)
            assert O.rel_err(ket.rdm(st, brawfn=bra), g[f"{tag}_t{k}"]) < TOL, (tag, st)
            assert O.rel_err(ket.expectationValue(st, brawfn=bra), g[f"{tag}_e{k}"]) < TOL
    # digits: one element, through the individual n-body path; equals the tensor's entry
    n, sz, norb = [int(x) for x in g["wa_meta"]]
    ket = fqe_b200.Wavefunction([[n, sz, norb]])
    ket.set_wfn(strategy="from_data", raw_data={(n, sz): g["wa_ket"]})
    r1 = ket.rdm("i^ j")
    # spin orbitals 2p (alpha) and 2p+1 (beta): <a+_{1a} a_{2a}> + <a+_{1b} a_{2b}> = rdm1[1,2]
    elem = ket.rdm("2^ 4") + ket.rdm("3^ 5")
    assert abs(elem - r1[1, 2]) < 1e-12
    with pytest.raises(TypeError):
        ket.rdm("i^ J")
    with pytest.raises(TypeError):
        ket.expectationValue(3.0)


def test_gram_kernel_against_torch():
    """fqeb_gram_accumulate (split-K DMMA Gram product) against torch.matmul in FP64 on ragged
    shapes, with and without the extra ket row, accumulating into a non-zero G"""
    import torch
    from fqe_b200 import lib as L
    lib = L.load()
    gen = torch.Generator(device="cuda").manual_seed(11)

    def rnd(*shape):
        return torch.complex(torch.randn(*shape, generator=gen, device="cuda", dtype=torch.float64),
                             torch.randn(*shape, generator=gen, device="cuda", dtype=torch.float64))

    for (m, n, ncols, ld, extra) in [(1, 1, 1, 1, False), (5, 7, 33, 40, False),
                                     (64, 65, 1000, 1000, True), (100, 101, 4097, 4100, True),
                                     (256, 257, 20011, 20011, True), (9, 1, 513, 600, True)]:
        bra = rnd(m, ld)
        nk = n - 1 if extra else n
        ket = rnd(max(nk, 1), ld)
        last = rnd(ncols) if extra else None
        g0 = rnd(m, n)
        got = g0.clone()
        L.call("fqeb_gram_accumulate", m, n, ncols, bra.data_ptr(), ld,
               ket.data_ptr() if nk > 0 else None, ld,
               last.data_ptr() if extra else None, got.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
        full = ket[:nk, :ncols] if nk > 0 else ket[:0, :ncols]
        if extra:
            full = torch.cat([full, last[None, :]], dim=0)
        ref = g0 + bra[:, :ncols].conj() @ full.transpose(0, 1)
        err = (got - ref).abs().max().item() / max(1.0, ref.abs().max().item())
        assert err < 1e-13, (m, n, ncols, err)
        if nk == m:
            # bra and ket the same buffer: the Hermitian shortcut (blocks below the diagonal are
            # mirrored instead of computed) must give the same matrix
            h = g0.clone()
            L.call("fqeb_gram_accumulate", m, n, ncols, bra.data_ptr(), ld, bra.data_ptr(), ld,
                   last.data_ptr() if extra else None, h.data_ptr(),
                   torch.cuda.current_stream().cuda_stream)
            fullh = bra[:, :ncols]
            if extra:
                fullh = torch.cat([fullh, last[None, :]], dim=0)
            refh = g0 + bra[:, :ncols].conj() @ fullh.transpose(0, 1)
            errh = (h - refh).abs().max().item() / max(1.0, refh.abs().max().item())
            assert errh < 1e-13, (m, n, ncols, errh)
        # bitwise reproducible
        again = g0.clone()
        L.call("fqeb_gram_accumulate", m, n, ncols, bra.data_ptr(), ld,
               ket.data_ptr() if nk > 0 else None, ld,
               last.data_ptr() if extra else None, again.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
        assert torch.equal(again, got)
