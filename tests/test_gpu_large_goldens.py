"""Parity with the UNMODIFIED reference at the sizes BASELINE.json's configs name.

tests/golden/ref_large.npz holds signatures (norm, 4096 sampled coefficients, every row / column
norm, 16 random rank-1 projections) of states the reference itself produced in the build
container (tests/golden/make_golden_large.py):

  sigma at (7,7,14) and (8,8,16), real 8-fold and complex-Hermitian integrals  FqeData.apply, C lm path
                                                   (reference fqe_data.py:685-710, lib/fqe_data.c:668-996)
  Taylor time_evolve at (7,7,14)                   wavefunction.py:548-568, 961-1054
  DiagonalCoulomb apply / evolve at (8,8,16), non-symmetric v     lib/fqe_data.c:455-602
  dense 3-body apply at norb=10, profile_3_body.py's tensor       fqe_data.py:1166-1216
  Chebyshev propagator, transform, individual 3-body operators at norb=12; rdm12 at norb=10, 12

Every sigma case runs on the default path AND on each alternative contraction path
(FQEB_FUSION=0: gather -> DMMA GEMM -> scatter;  FQEB_OZAKI=0: FP64 DMMA instead of the
INT8-sliced tensor-core contraction).  Tolerance: 1e-10 relative (BASELINE.json north_star).
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_large as GL  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def large(golden_dir):
    path = os.path.join(golden_dir, "ref_large.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def _check(large, tag, tensor):
    sig = GL.stored_signature(large, tag)
    got = GL.signature_torch(tensor, int(sig["seed"][0]))
    worst, errs = GL.signature_error(sig, got)
    assert worst < TOL, f"{tag}: {errs}"
    return worst


class _Env:
    def __init__(self, **kv):
        self.kv, self.old = kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


PATHS = {
    "default": {},
    "dmma_fused": {"FQEB_OZAKI": "0"},
    "three_kernel": {"FQEB_OZAKI": "0", "FQEB_FUSION": "0"},
}


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("kind", ["real8", "herm"])
@pytest.mark.parametrize("norb", [14, 16])
def test_sigma_matches_reference_at_config_size(large, norb, kind, path):
    import fqe_b200 as fqe
    from fqe_b200 import synth
    from fqe_b200.fqe_data import release_workspace
    tag = f"sigma{norb}_{kind}"
    if f"{tag}_norm" not in large:
        pytest.skip(f"{tag} not in ref_large.npz")
    if kind == "herm" and path == "dmma_fused":
        pytest.skip("complex operators have one DMMA path (three_kernel)")
    n, sz = norb, 0
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    h1, h2 = synth.integrals(norb, kind)
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data",
                raw_data={(n, sz): synth.state(la, lb, seed=synth.seed_for(norb, 50))})
    with _Env(**PATHS[path]):
        out = wfn.sector((n, sz)).apply((h1, h2))
        torch.cuda.synchronize()
    _check(large, tag, out.coeff)
    del out, wfn
    release_workspace()


def test_taylor_time_evolve_norb14(large):
    import fqe_b200 as fqe
    from fqe_b200 import synth
    from fqe_b200.fqe_data import release_workspace
    if "taylor14_norm" not in large:
        pytest.skip("taylor14 not in ref_large.npz")
    n, sz, norb = [int(x) for x in large["taylor14_meta"]]
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    h1, h2 = synth.integrals(norb, "real8", scale=float(large["taylor14_scale"][0]))
    ham = fqe.get_restricted_hamiltonian((h1, h2), e_0=float(large["taylor14_e0"][0]))
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data",
                raw_data={(n, sz): synth.state(la, lb, seed=int(large["taylor14_state_seed"][0]))})
    t = float(large["taylor14_t"][0])
    for path in ("default", "dmma_fused"):
        with _Env(**PATHS[path]):
            ev = wfn.time_evolve(t, ham)
            _check(large, "taylor14", ev.get_coeff_device((n, sz)))
            agu = wfn.apply_generated_unitary(t, "taylor", ham)
            _check(large, "taylor14_agu", agu.get_coeff_device((n, sz)))
            del ev, agu
    release_workspace()


def test_diagonal_coulomb_norb16(large):
    import fqe_b200 as fqe
    from fqe_b200 import synth
    if "dc16_apply_norm" not in large:
        pytest.skip("dc16 not in ref_large.npz")
    n, sz, norb = [int(x) for x in large["dc16_apply_meta"]]
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data",
                raw_data={(n, sz): synth.state(la, lb, seed=synth.seed_for(norb, 50))})
    vij = synth.diagonal_coulomb_matrix(norb, int(large["dc16_apply_vij_seed"][0]),
                                        symmetric=False)
    e0, t = float(large["dc16_apply_e0"][0]), float(large["dc16_apply_t"][0])
    dch = fqe.get_diagonalcoulomb_hamiltonian(vij, e_0=e0)
    _check(large, "dc16_apply", wfn.apply(dch).get_coeff_device((n, sz)))
    _check(large, "dc16_evolve", wfn.time_evolve(t, dch).get_coeff_device((n, sz)))
    h4 = np.zeros((norb,) * 4)
    for i in range(norb):
        for j in range(norb):
            h4[i, j, i, j] = -vij[i, j]
        h4[i, i, i, i] = large["dc16_h4diag"][i]
    out = wfn.apply(fqe.get_diagonalcoulomb_hamiltonian(h4))
    _check(large, "dc16_apply4", out.get_coeff_device((n, sz)))


@pytest.mark.parametrize("norb", [10, 12])
def test_dense_three_body_profile_shape(large, norb):
    import fqe_b200 as fqe
    from fqe_b200 import synth
    from fqe_b200.fqe_data import release_workspace
    tag = f"body3_{norb}"
    if f"{tag}_norm" not in large:
        pytest.skip(f"{tag} not in ref_large.npz")
    n, sz = norb, 0
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    idx = np.indices((norb,) * 6)
    h3 = ((idx[0] + idx[3]) * (idx[1] + idx[4]) * (idx[2] + idx[5]) * 0.002).astype(np.complex128)
    del idx
    h1 = np.zeros((norb,) * 2, dtype=np.complex128)
    h2 = np.zeros((norb,) * 4, dtype=np.complex128)
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data",
                raw_data={(n, sz): synth.state(la, lb, seed=synth.seed_for(norb, 53))})
    out = wfn.apply((h1, h2, h3))
    _check(large, tag, out.get_coeff_device((n, sz)))
    del out, wfn
    release_workspace()


@pytest.mark.parametrize("norb", [10, 12])
def test_rdm12_at_scale(large, norb):
    """FqeData.rdm12, plain and transition, against the reference's own output at norb = 10 and
    12 (reference fqe_data.py:1726-1838): rdm1 in full, rdm2 through its norm, its ijij trace and
    4096 sampled entries; the reductions run on the library's Gram kernel (Hermitian shortcut for
    the plain form, general form for the transition one)."""
    import fqe_b200 as fqe
    from fqe_b200 import synth
    if f"rdm{norb}_meta" not in large:
        pytest.skip(f"rdm{norb} not in ref_large.npz")
    n, sz = norb, 0
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    ket = fqe.Wavefunction([[n, sz, norb]])
    ket.set_wfn(strategy="from_data",
                raw_data={(n, sz): synth.state(la, lb, seed=synth.seed_for(norb, 54))})
    bra = fqe.Wavefunction([[n, sz, norb]])
    bra.set_wfn(strategy="from_data",
                raw_data={(n, sz): synth.state(la, lb, seed=synth.seed_for(norb, 55))})
    ksec, bsec = ket.sector((n, sz)), bra.sector((n, sz))
    for tag, (g1, g2) in ((f"rdm{norb}", ksec.rdm12()), (f"trdm{norb}", ksec.rdm12(bsec))):
        r1 = large[f"{tag}_1"]
        assert np.abs(g1 - r1).max() < 1e-11 * np.abs(r1).max(), tag
        flat = g2.reshape(-1)
        scale = float(large[f"{tag}_2_norm"][0])
        assert abs(np.linalg.norm(g2) - scale) < 1e-11 * scale, tag
        assert np.abs(flat[large[f"{tag}_2_idx"]] - large[f"{tag}_2_val"]).max() < \
            1e-11 * np.abs(large[f"{tag}_2_val"]).max(), tag
        assert abs(np.einsum("ijij", g2) - large[f"{tag}_2_trace"][0]) < 1e-10 * scale, tag


def test_chebyshev_and_transform_norb12(large):
    """apply_generated_unitary(algo='chebyshev') (reference wavefunction.py:570-611) and
    Wavefunction.transform (:813-959) at norb = 12 against the reference's own outputs"""
    import fqe_b200 as fqe
    from fqe_b200 import synth
    from fqe_b200.fqe_data import release_workspace
    if "cheb12_norm" not in large:
        pytest.skip("cheb12 not in ref_large.npz")
    n, sz, norb = [int(x) for x in large["cheb12_meta"]]
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    h1, h2 = synth.integrals(norb, "real8", scale=float(large["cheb12_scale"][0]))
    ham = fqe.get_restricted_hamiltonian((h1, h2), e_0=float(large["cheb12_e0"][0]))
    c0 = synth.state(la, lb, seed=synth.seed_for(norb, 56))
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    t, lim = float(large["cheb12_t"][0]), [float(x) for x in large["cheb12_spec_lim"]]
    for path in ("default", "dmma_fused"):
        with _Env(**PATHS[path]):
            ch = wfn.apply_generated_unitary(t, "chebyshev", ham, spec_lim=lim)
            _check(large, "cheb12", ch.get_coeff_device((n, sz)))
            del ch
    w2 = fqe.Wavefunction([[n, sz, norb]])
    w2.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    w2.transform(large["transform12_rot"])
    _check(large, "transform12", w2.get_coeff_device((n, sz)))
    release_workspace()


def test_individual_nbody_norb12(large):
    """individual n-body operators at norb = 12 (BASELINE config 5's sweep shape): apply and exact
    evolution (reference fqe_data.py:1558-1653, 2385-2590) against the reference's own outputs"""
    import ast
    import fqe_b200 as fqe
    from fqe_b200 import synth
    if "nbody12_meta" not in large:
        pytest.skip("nbody12 not in ref_large.npz")
    n, sz, norb = [int(x) for x in large["nbody12_meta"]]
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    c0 = synth.state(la, lb, seed=synth.seed_for(norb, 57))
    wfn = fqe.Wavefunction([[n, sz, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
    sec = wfn.sector((n, sz))
    zc, tm = complex(large["nbody12_coeff"][0]), float(large["nbody12_time"][0])
    for k, o in enumerate(large["nbody12_ops"]):
        da, ua, db, ub = ast.literal_eval(str(o))
        out = sec.apply_individual_nbody(zc, da, ua, db, ub)
        _check(large, f"nbody12_apply{k}", out.coeff)
        if da == ua and db == ub:
            tmp = fqe.Wavefunction([[n, sz, norb]])
            tmp.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
            tmp.sector((n, sz)).evolve_inplace_individual_nbody_trivial(tm, zc, da, db)
            _check(large, f"nbody12_evolve{k}", tmp.sector((n, sz)).coeff)
        else:
            ev = sec.evolve_individual_nbody_nontrivial(tm, zc, da, ua, db, ub)
            _check(large, f"nbody12_evolve{k}", ev.coeff)
