"""Per-kernel parity through the C ABI: gather (make_dvec), scatter (make_coeff),
DMMA contraction, diagonal Coulomb and BLAS-1, each against the CPU oracle on the
same seeded inputs.

Tolerances: the gather only multiplies by +-1, so D is compared BIT-EXACTLY;
floating-point kernels are compared at 1e-12 relative 2-norm (target 1e-10)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import fqe_oracle as O
from oracle import ref_harness as R

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _rand_c(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def _data(na, nb, norb, seed=0):
    from fqe_b200.fqe_data import FqeData
    d = FqeData(na, nb, norb)
    rng = np.random.default_rng(1000 * norb + 10 * na + nb + seed)
    c = _rand_c(rng, (d.lena(), d.lenb()))
    d.set_wfn(strategy="from_data", raw_data=c)
    return d, c, rng


CFGS = [(1, 1, 2), (2, 1, 4), (2, 3, 6), (3, 3, 6), (4, 3, 7), (4, 4, 8), (0, 3, 5), (5, 0, 5),
        (5, 5, 10), (2, 6, 9)]


@pytest.mark.parametrize("cfg", CFGS)
def test_make_dvec_bit_exact(cfg):
    na, nb, norb = cfg
    d, c, _ = _data(na, nb, norb)
    dvec = d.calculate_dvec_spatial().cpu().numpy()
    ref = O.dvec_spatial(O.graph(na, nb, norb), c)
    assert dvec.shape == ref.shape
    assert np.array_equal(dvec, ref)


@pytest.mark.parametrize("cfg", CFGS)
def test_make_coeff(cfg):
    na, nb, norb = cfg
    d, c, rng = _data(na, nb, norb)
    e = _rand_c(rng, (norb, norb, d.lena(), d.lenb()))
    out = d._calculate_coeff_spatial_with_dvec(torch.from_numpy(e).cuda()).cpu().numpy()
    ref = O.coeff_from_dvec(O.graph(na, nb, norb), e)
    assert O.rel_err(out, ref) < TOL


def test_make_dvec_chunks_and_slices():
    """row chunks and pair slices tile the full tensor exactly"""
    from fqe_b200 import lib as L
    na, nb, norb = 3, 4, 7
    d, c, _ = _data(na, nb, norb)
    full = d.calculate_dvec_spatial()
    npair = norb * norb
    la, lb = d.lena(), d.lenb()
    for (r0, nr, p0, p1) in [(0, 5, 0, npair), (5, la - 5, 10, 30), (la - 1, 1, 48, 49)]:
        ldd = nr * lb + 3
        buf = torch.full((p1 - p0, ldd), 7.0 + 7.0j, dtype=torch.complex128, device="cuda")
        L.call("fqeb_make_dvec", d._core.handle, d.coeff.data_ptr(), buf.data_ptr(), ldd, r0, nr,
               p0, p1, None)
        torch.cuda.synchronize()
        got = buf[:, :nr * lb].reshape(p1 - p0, nr, lb)
        exp = full.reshape(npair, la, lb)[p0:p1, r0:r0 + nr]
        assert torch.equal(got, exp)
        assert torch.all(buf[:, nr * lb:] == 7.0 + 7.0j)  # padding untouched


def test_make_coeff_chunks_sum_to_full():
    from fqe_b200 import lib as L
    na, nb, norb = 4, 3, 7
    d, c, rng = _data(na, nb, norb)
    la, lb = d.lena(), d.lenb()
    npair = norb * norb
    e = torch.from_numpy(_rand_c(rng, (npair, la, lb))).cuda()
    full = d._calculate_coeff_spatial_with_dvec(e.reshape(norb, norb, la, lb))
    acc = torch.zeros((la, lb), dtype=torch.complex128, device="cuda")
    z = 0.3 - 1.1j
    for r0, nr in [(0, 9), (9, 20), (29, la - 29)]:
        chunk = e[:, r0:r0 + nr].contiguous()
        L.call("fqeb_make_coeff", d._core.handle, chunk.data_ptr(), nr * lb, r0, nr, z.real,
               z.imag, acc.data_ptr(), None)
    torch.cuda.synchronize()
    assert O.rel_err(acc.cpu().numpy(), z * full.cpu().numpy()) < TOL


@pytest.mark.parametrize("norb", [2, 3, 4, 6, 7, 8, 12, 14, 16])
@pytest.mark.parametrize("kind", ["real", "imag", "complex"])
def test_contract_matches_matmul(norb, kind):
    """DMMA GEMM against torch's FP64 matmul (floating-point kernel: plain library
    reference of the same op)."""
    from fqe_b200 import lib as L
    from fqe_b200.fqe_data import DenseOperator
    rng = np.random.default_rng(norb * 7 + len(kind))
    npair = norb * norb
    h2p = _rand_c(rng, (npair, npair))
    if kind == "real":
        h2p = h2p.real.astype(np.complex128)
    elif kind == "imag":
        h2p = 1j * h2p.imag
    # DenseOperator folds h2 -> h2p = -moveaxis(h2,1,2); invert that to feed h2p exactly
    h2p4 = h2p.reshape((norb,) * 4)
    h2 = -np.moveaxis(h2p4, 2, 1)
    op = DenseOperator(norb, np.zeros((norb, norb)), h2)
    assert op.kind == {"real": L.OP_REAL, "imag": L.OP_IMAG, "complex": L.OP_COMPLEX}[kind]
    assert np.array_equal(op._h2p.reshape(npair, npair), h2p)
    lib = L.load()
    for ncols, (p0, p1) in [(1, (0, npair)), (300, (0, npair)), (129, (2, min(npair, 11)))]:
        ld = ((ncols + 127) // 128) * 128
        nij = p1 - p0
        drows = lib.fqeb_contract_dvec_rows(op.handle, nij)
        dv = torch.zeros((drows, ld), dtype=torch.complex128, device="cuda")
        dhost = _rand_c(rng, (nij, ncols))
        dv[:nij, :ncols] = torch.from_numpy(dhost).cuda()
        ev = torch.full((npair + 8, ld), 5.0, dtype=torch.complex128, device="cuda")
        L.call("fqeb_contract", op.handle, dv.data_ptr(), ld, ev.data_ptr(), ld, ncols, p0, p1,
               None)
        torch.cuda.synchronize()
        a = h2p[:, p0:p1]
        if kind == "imag":
            a = a / 1j  # the library factors i out (applied later by the scatter)
        ref = (torch.from_numpy(a).cuda() @ torch.from_numpy(dhost).cuda()).cpu().numpy()
        got = ev[:npair, :ncols].cpu().numpy()
        assert O.rel_err(got, ref) < TOL, (ncols, p0, p1)
        assert torch.all(ev[npair:] == 5.0)  # rows beyond norb^2 untouched


def test_contract_pair_symmetric_operator():
    """h2'[ij,kl] symmetric under i<->j, k<->l -> contraction in the i>=j pair space"""
    from fqe_b200 import lib as L
    from fqe_b200.fqe_data import DenseOperator
    for norb, cplx in [(3, False), (6, True), (8, False), (16, False), (16, True)]:
        rng = np.random.default_rng(norb + 5)
        w = _rand_c(rng, (norb,) * 4) if cplx else rng.standard_normal((norb,) * 4) + 0j
        h2p4 = w + w.transpose(1, 0, 2, 3)
        h2p4 = h2p4 + h2p4.transpose(0, 1, 3, 2)
        op = DenseOperator(norb, np.zeros((norb, norb)), -np.moveaxis(h2p4, 2, 1))
        npc = norb * (norb + 1) // 2
        assert op.symmetric and op.npair == npc
        assert op.kind == (L.OP_COMPLEX if cplx else L.OP_REAL)
        pairs = [(i, j) for i in range(norb) for j in range(i + 1)]
        h2c = np.array([[h2p4[i, j, k, l] for (k, l) in pairs] for (i, j) in pairs])
        lib = L.load()
        ncols = 200
        ld = 256
        drows = lib.fqeb_contract_dvec_rows(op.handle, npc)
        dv = torch.zeros((drows, ld), dtype=torch.complex128, device="cuda")
        dhost = _rand_c(rng, (npc, ncols))
        dv[:npc, :ncols] = torch.from_numpy(dhost).cuda()
        ev = torch.zeros((npc + 8, ld), dtype=torch.complex128, device="cuda")
        L.call("fqeb_contract", op.handle, dv.data_ptr(), ld, ev.data_ptr(), ld, ncols, 0, npc,
               None)
        torch.cuda.synchronize()
        assert O.rel_err(ev[:npc, :ncols].cpu().numpy(), h2c @ dhost) < TOL


def test_symmetry_switch(monkeypatch, contraction):
    """FQEB_NO_SYMMETRY disables the compression; both routes give the same sigma (on either
    contraction back end, each held to its own bound)"""
    from conftest import sigma_tol
    from fqe_b200 import synth
    TOL = sigma_tol()
    from fqe_b200.fqe_data import DenseOperator, FqeData
    na, nb, norb = 3, 4, 7
    h1, h2 = synth.integrals(norb, "real8")
    d, c, _ = _data(na, nb, norb)
    op_sym = DenseOperator(norb, h1, h2)
    assert op_sym.symmetric and op_sym.npair == 28
    op_t = DenseOperator(norb, -0.1j * h1, -0.1j * h2)   # Taylor's iht tensors
    assert op_t.symmetric and op_t.kind == 1
    monkeypatch.setenv("FQEB_NO_SYMMETRY", "1")
    op_full = DenseOperator(norb, h1, h2)
    assert (not op_full.symmetric) and op_full.npair == 49
    ref = O.sigma_restricted(O.graph(na, nb, norb), c, h1, h2)
    assert O.rel_err(d.apply_operator(op_sym).cpu().numpy(), ref) < TOL
    assert O.rel_err(d.apply_operator(op_full).cpu().numpy(), ref) < TOL
    assert O.rel_err(d.apply_operator(op_t).cpu().numpy(), -0.1j * ref) < TOL


def test_contraction_path_follows_the_state(monkeypatch):
    """The INT8-sliced contraction quantises C with ONE global scale; the library estimates the
    error from max |C|, ||C|| and the number of non-zero parts and falls back to the FP64 DMMA
    kernel when it is above FQEB_OZAKI_TOL.  A single determinant and a dense random state take
    the tensor-core path, a dense state with one dominant coefficient the FP64 one; all agree
    with the oracle."""
    from fqe_b200 import lib as L, synth
    from fqe_b200.fqe_data import DenseOperator, FqeData
    monkeypatch.delenv("FQEB_OZAKI", raising=False)
    monkeypatch.delenv("FQEB_FUSION", raising=False)
    lib = L.load()
    na, nb, norb = 4, 4, 8
    h1, h2 = synth.integrals(norb, "real8")
    op = DenseOperator(norb, h1, h2)
    g = O.graph(na, nb, norb)
    d = FqeData(na, nb, norb)
    rng = np.random.default_rng(77)
    dense = _rand_c(rng, (d.lena(), d.lenb()))
    hf = np.zeros_like(dense)
    hf[0, 0] = 1.0
    peaked = 1e-7 * dense
    peaked[3, 5] = 1.0
    for state, want, tol in [(dense, 3, 2e-11), (hf, 3, 2e-11), (peaked, 2, 1e-12)]:
        d.set_wfn(strategy="from_data", raw_data=state)
        out = d.apply_operator(op).cpu().numpy()
        assert lib.fqeb_sigma_last_path() == want
        assert O.rel_err(out, O.sigma_restricted(g, state, h1, h2)) < tol
    # a zero vector: sigma is zero on either path
    d.set_wfn(strategy="from_data", raw_data=np.zeros_like(dense))
    assert not d.apply_operator(op).any()
    # NaN in the state is propagated (FP64 path), not silently replaced by zeros
    bad = dense.copy()
    bad[2, 2] = np.nan
    d.set_wfn(strategy="from_data", raw_data=bad)
    out = d.apply_operator(op)
    assert lib.fqeb_sigma_last_path() == 2 and bool(torch.isnan(torch.view_as_real(out)).any())


DC_CFGS = [(2, 3, 6), (2, 1, 4), (4, 4, 8), (0, 2, 4), (3, 3, 3), (5, 4, 9)]


@pytest.mark.parametrize("cfg", DC_CFGS)
def test_dc_apply_and_evolve(cfg):
    na, nb, norb = cfg
    d, c, rng = _data(na, nb, norb)
    g = O.graph(na, nb, norb)
    diag = _rand_c(rng, norb)
    v = _rand_c(rng, (norb, norb))  # deliberately non-symmetric
    out = d.apply_diagonal_coulomb(diag, v).cpu().numpy()
    assert O.rel_err(out, O.dc_apply(g, c, diag, v)) < TOL
    assert np.array_equal(d.to_numpy(), c)  # out of place
    out = d.evolve_diagonal_coulomb(0.1 * diag, 0.1 * v).cpu().numpy()
    assert O.rel_err(out, O.dc_evolve(g, c, 0.1 * diag, 0.1 * v)) < TOL
    if R.available():
        rg = R.graph(na, nb, norb)
        assert O.rel_err(out, R.dc_evolve(rg, c, 0.1 * diag, 0.1 * v)) < TOL
    d.evolve_diagonal_coulomb(0.1 * diag, 0.1 * v, inplace=True)
    assert np.array_equal(d.to_numpy(), out)


def test_dc_shipped_golden(golden_dir):
    """reference tests/fqe_data_test.py:268-280 (evolve, t=0.1, non-symmetric dmat)"""
    shipped = np.load(os.path.join(golden_dir, "ref_unittest_fqe_data.npz"))
    from fqe_b200.fqe_data import FqeData
    for (na, nb, norb) in [(2, 3, 6), (2, 1, 4)]:
        s = f"{na:02d}{nb:02d}{norb:02d}"
        d = FqeData(na, nb, norb)
        shp = (d.lena(), d.lenb())
        c = (shipped["cr" + s] + 1j * shipped["ci" + s]).reshape(shp)
        d.set_wfn(strategy="from_data", raw_data=c)
        dmat = shipped["dmat" + s].reshape(norb, norb)
        ref = (shipped[f"cr{s}_dc"] + 1j * shipped[f"ci{s}_dc"]).reshape(shp)
        out = d.evolve_diagonal_coulomb(np.zeros(norb), -0.1j * dmat).cpu().numpy()
        assert O.rel_err(out, ref) < TOL


def test_blas1():
    from fqe_b200.fqe_data import FqeData
    for (na, nb, norb) in [(1, 1, 2), (3, 3, 6), (6, 6, 12)]:
        x, cx, rng = _data(na, nb, norb, seed=1)
        y, cy, _ = _data(na, nb, norb, seed=2)
        a = 0.7 - 0.2j
        assert abs(x.norm() - np.linalg.norm(cx)) < 1e-12 * np.linalg.norm(cx)
        assert abs(x.vdot(y) - np.vdot(cx, cy)) < 1e-12 * np.linalg.norm(cx) * np.linalg.norm(cy)
        y.ax_plus_y(a, x)
        assert O.rel_err(y.to_numpy(), cy + a * cx) < 1e-15
        nrm = y.axpy_norm(a, x)
        assert abs(nrm - np.linalg.norm(cx)) < 1e-12 * np.linalg.norm(cx)
        assert O.rel_err(y.to_numpy(), cy + 2 * a * cx) < 1e-15
        x.scale(a)
        assert O.rel_err(x.to_numpy(), a * cx) < 1e-15
        # determinism of the two-pass reduction
        assert x.norm() == x.norm()
