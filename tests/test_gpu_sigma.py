"""sigma = H C through the public API and the C ABI against the oracle, the
reference's golden vectors and the compiled reference C path.

Tolerance (BASELINE.json north_star): relative 2-norm <= 1e-10 in complex128; the tests
run every case on both contraction back ends (``contraction`` fixture, tests/conftest.py) and
assert 1e-12 on the FP64 DMMA kernels and 2e-11 on the INT8-sliced tensor-core kernel."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import fqe_oracle as O
from oracle import ref_harness as R

from conftest import sigma_tol as TOL   # 1e-12 on the FP64 paths, 2e-11 on the sliced one

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("contraction")]


def _wfn(n, sz, norb, c):
    import fqe_b200
    w = fqe_b200.Wavefunction([[n, sz, norb]])
    w.set_wfn(strategy="from_data", raw_data={(n, sz): c})
    return w


def _case(na, nb, norb, kind, seed=0):
    from fqe_b200 import synth
    h1, h2 = synth.integrals(norb, kind, seed=synth.seed_for(norb, seed))
    g = O.graph(na, nb, norb)
    c = synth.state(g.lena, g.lenb, seed=synth.seed_for(norb, seed + 50))
    return g, c, h1, h2


SIGMA_CFGS = [(1, 1, 2), (2, 1, 4), (2, 2, 4), (2, 3, 6), (3, 3, 6), (4, 3, 7), (4, 4, 8),
              (0, 2, 4), (3, 0, 5), (4, 4, 4), (1, 6, 8), (5, 5, 10)]


@pytest.mark.parametrize("cfg", SIGMA_CFGS)
@pytest.mark.parametrize("kind", ["real8", "herm", "general"])
def test_sigma_vs_oracle(cfg, kind):
    na, nb, norb = cfg
    g, c, h1, h2 = _case(na, nb, norb, kind)
    from fqe_b200.fqe_data import FqeData
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    out = d.apply((h1, h2))
    ref = O.sigma_restricted(g, c, h1, h2)
    assert O.rel_err(out.to_numpy(), ref) < TOL()
    assert np.array_equal(d.to_numpy(), c)  # apply is out of place
    # purely imaginary operator (the Taylor iht tensors) takes the real-GEMM route
    out = d.apply((-0.05j * h1, -0.05j * h2))
    assert O.rel_err(out.to_numpy(), -0.05j * ref) < TOL()
    # one-body only
    out = d.apply((h1,))
    assert O.rel_err(out.to_numpy(), O.sigma_one_body(g, c, h1)) < TOL()


def test_sigma_real_dtype_inputs():
    """float64 tensors are accepted like the reference (pyx wrappers cast to c128)"""
    na, nb, norb = 3, 3, 6
    g, c, h1, h2 = _case(na, nb, norb, "real8")
    from fqe_b200.fqe_data import FqeData
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    out = d.apply((h1.real.copy(), h2.real.copy()))
    assert O.rel_err(out.to_numpy(), O.sigma_restricted(g, c, h1, h2)) < TOL()


@pytest.mark.parametrize("cfg", [(2, 3, 6), (2, 1, 4), (1, 1, 2)])
def test_sigma_shipped_goldens(golden_dir, cfg):
    """reference tests/fqe_data_test.py:443-456, 522-565 (golden _1, _2, _12)"""
    shipped = np.load(os.path.join(golden_dir, "ref_unittest_fqe_data.npz"))
    na, nb, norb = cfg
    s = f"{na:02d}{nb:02d}{norb:02d}"
    from fqe_b200.fqe_data import FqeData
    d = FqeData(na, nb, norb)
    shp = (d.lena(), d.lenb())
    c = (shipped["cr" + s] + 1j * shipped["ci" + s]).reshape(shp)
    d.set_wfn(strategy="from_data", raw_data=c)
    h1 = shipped["h1" + s].reshape((norb,) * 2)
    h2 = shipped["h2" + s].reshape((norb,) * 4)

    def ref(tag):
        return (shipped[f"cr{s}_{tag}"] + 1j * shipped[f"ci{s}_{tag}"]).reshape(shp)

    assert O.rel_err(d.apply((h1, h2)).to_numpy(), ref("12")) < TOL()
    assert O.rel_err(d.apply((np.zeros_like(h1), h2)).to_numpy(), ref("2")) < TOL()
    assert O.rel_err(d.apply((h1,)).to_numpy(), ref("1")) < TOL()


def test_sigma_chunked_and_sharded():
    """small workspaces force many alpha-row chunks; row / pair shards sum to sigma"""
    from fqe_b200 import lib as L
    from fqe_b200.fqe_data import DenseOperator, FqeData
    na, nb, norb = 4, 4, 8
    for kind in ("real8", "herm"):
        g, c, h1, h2 = _case(na, nb, norb, kind)
        ref = O.sigma_restricted(g, c, h1, h2)
        d = FqeData(na, nb, norb)
        d.set_wfn(strategy="from_data", raw_data=c)
        op = DenseOperator(norb, h1, h2)
        lib = L.load()
        npair = op.npair
        la = d.lena()

        def run(rows_per_chunk, r0, r1, p0, p1):
            nbytes = int(lib.fqeb_sigma_workspace_bytes(d._core.handle, op.handle,
                                                        rows_per_chunk, p0, p1))
            ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            out = torch.empty_like(d.coeff)
            L.call("fqeb_sigma_restricted", d._core.handle, op.handle, d.coeff.data_ptr(),
                   out.data_ptr(), ws.data_ptr(), nbytes, r0, r1, p0, p1, None)
            torch.cuda.synchronize()
            return out.cpu().numpy()

        for rows in (1, 3, 17, la):
            assert O.rel_err(run(rows, 0, la, 0, npair), ref) < TOL(), rows
        # determinant-row shards (world of 3)
        parts = [run(7, r0, r1, 0, npair) for r0, r1 in [(0, 20), (20, 50), (50, la)]]
        assert O.rel_err(sum(parts), ref) < TOL()
        # pair (ij) shards as in north_star (world of 4)
        from fqe_b200.distributed import shard_plan
        parts = [run(11, *shard_plan("pair", r, 4, la, npair)[0], *shard_plan("pair", r, 4, la, npair)[1])
                 for r in range(4)]
        assert O.rel_err(sum(parts), ref) < TOL()
        parts = [run(11, 0, la, p0, p1) for p0, p1 in [(0, 6), (6, 20), (20, npair)]]
        assert O.rel_err(sum(parts), ref) < TOL()
        # too-small workspace is an error, not a crash
        out = torch.empty_like(d.coeff)
        ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
        rc = lib.fqeb_sigma_restricted(d._core.handle, op.handle, d.coeff.data_ptr(),
                                       out.data_ptr(), ws.data_ptr(), 1024, 0, la, 0, npair, None)
        assert rc == L.ERR_NOMEM


def test_sigma_host_entry():
    """the reference-facing C-ABI call with HOST buffers"""
    from fqe_b200 import lib as L
    from fqe_b200.fqe_data import DenseOperator
    na, nb, norb = 3, 4, 7
    g, c, h1, h2 = _case(na, nb, norb, "herm")
    h1p, h2p = O.fold_restricted(h1, h2)
    h1p = np.ascontiguousarray(h1p)
    h2p = np.ascontiguousarray(h2p)
    out = np.zeros_like(c)
    cc = np.ascontiguousarray(c)
    L.call("fqeb_sigma_restricted_host", norb, na, nb, h1p.ctypes.data, h2p.ctypes.data,
           cc.ctypes.data, out.ctypes.data)
    assert O.rel_err(out, O.sigma_restricted(g, c, h1, h2)) < TOL()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg,kind", [((5, 5, 10), "real8"), ((6, 6, 12), "herm"),
                                      ((6, 5, 11), "general"), ((6, 6, 12), "real8")])
def test_sigma_vs_compiled_reference(cfg, kind):
    """the production C `lm` path of the reference, compiled from its sources"""
    na, nb, norb = cfg
    from fqe_b200 import synth
    from fqe_b200.fqe_data import FqeData
    h1, h2 = synth.integrals(norb, kind)
    rg = R.graph(na, nb, norb)
    c = synth.state(rg.lena, rg.lenb, seed=5 + norb)
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    out = d.apply((h1, h2)).to_numpy()
    assert O.rel_err(out, R.sigma_restricted(rg, c, h1, h2)) < TOL()


def test_properties_at_full_size():
    """size-independent checks at norb=14 (11.8M determinants), where the oracle is
    too slow: linearity, Hermiticity <x|Hy> = <Hx|y>, the diagonal-Coulomb
    apply == dense apply identity (reference tests/fqe_data_test.py:297-325)"""
    import fqe_b200
    from fqe_b200 import synth
    norb, n, sz = 14, 14, 0
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    h1, h2 = synth.integrals(norb, "herm", scale=0.02)
    ham = fqe_b200.get_restricted_hamiltonian((h1, h2))
    x = fqe_b200.Wavefunction([[n, sz, norb]])
    y = fqe_b200.Wavefunction([[n, sz, norb]])
    gen = torch.Generator(device="cuda").manual_seed(1234)
    for w in (x, y):
        t = torch.randn((la, lb, 2), dtype=torch.float64, device="cuda", generator=gen)
        w.set_wfn(strategy="from_data", raw_data={(n, sz): torch.view_as_complex(t)})
        w.normalize()
    hx, hy = x.apply(ham), y.apply(ham)
    # Hermiticity
    lhs, rhs = x.vdot(hy), hx.vdot(y)
    assert abs(lhs - rhs) < 1e-11 * max(1.0, abs(lhs))
    # linearity
    a, b = 0.3 - 0.7j, -1.1 + 0.2j
    z = fqe_b200.Wavefunction([[n, sz, norb]])
    z.set_wfn(strategy="zero")
    z.ax_plus_y(a, x)
    z.ax_plus_y(b, y)
    hz = z.apply(ham)
    hz.ax_plus_y(-a, hx)
    hz.ax_plus_y(-b, hy)
    assert hz.norm() < 1e-12 * (hx.norm() + hy.norm())
    # diagonal Coulomb identity with a non-symmetric v
    rng = np.random.default_rng(454417)
    vij = 8 * rng.uniform(0, 1, size=(norb, norb)) / norb
    h2d = np.zeros((norb,) * 4)
    for i in range(norb):
        for j in range(norb):
            h2d[i, j, i, j] = -vij[i, j]
    dense = x.apply(fqe_b200.get_restricted_hamiltonian((np.zeros((norb, norb)), h2d)))
    diag = x.apply(fqe_b200.get_diagonalcoulomb_hamiltonian(h2d))
    dense.ax_plus_y(-1.0, diag)
    assert dense.norm() < TOL() * diag.norm()


@pytest.mark.parametrize("cfg", [(2, 2, 4), (3, 3, 6), (4, 3, 7), (4, 4, 8), (5, 5, 10), (6, 6, 12),
                                 (2, 5, 9), (8, 8, 16)])
def test_fused_gather_contraction_matches_three_kernel_path(cfg, monkeypatch):
    """FQEB_FUSION=1 routes real-orbital integrals through the fused gather+DMMA kernel (D
    never materialised); the default is gather -> GEMM -> scatter.  Both must agree with
    each other (and with the oracle where it is fast enough)."""
    from fqe_b200 import lib as L
    from fqe_b200 import synth
    from fqe_b200.fqe_data import DenseOperator, FqeData
    na, nb, norb = cfg
    h1, h2 = synth.integrals(norb, "real8")
    d = FqeData(na, nb, norb)
    c = synth.state(d.lena(), d.lenb(), seed=99 + norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    for scale in (1.0, -0.01j):          # real class and Taylor's imaginary class
        op = DenseOperator(norb, scale * h1, scale * h2)
        assert op.symmetric
        monkeypatch.setenv("FQEB_FUSION", "1")
        fused = d.apply_operator(op)
        monkeypatch.setenv("FQEB_FUSION", "0")
        plain = d.apply_operator(op)
        monkeypatch.setenv("FQEB_FUSION", "1")
        err = float(torch.linalg.norm(fused - plain) / torch.linalg.norm(plain))
        assert err < TOL(), (cfg, scale, err)
        if norb <= 10:
            ref = scale * O.sigma_restricted(O.graph(na, nb, norb), c, h1, h2)
            assert O.rel_err(fused.cpu().numpy(), ref) < TOL()
    # shards and small workspaces through the fused path
    lib = L.load()
    op = DenseOperator(norb, h1, h2)
    la, npair = d.lena(), op.npair
    full = d.apply_operator(op)

    def run(rows_per_chunk, r0, r1, p0, p1):
        nbytes = int(lib.fqeb_sigma_workspace_bytes(d._core.handle, op.handle, rows_per_chunk,
                                                    p0, p1))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device="cuda")
        out = torch.empty_like(d.coeff)
        L.call("fqeb_sigma_restricted", d._core.handle, op.handle, d.coeff.data_ptr(),
               out.data_ptr(), ws.data_ptr(), nbytes, r0, r1, p0, p1, None)
        return out

    if norb <= 12:
        acc = run(3, 0, la // 2, 0, npair) + run(5, la // 2, la, 0, npair)
        assert float(torch.linalg.norm(acc - full) / torch.linalg.norm(full)) < TOL()
        half = (npair // 2) & ~1
        acc = run(la, 0, la, 0, half) + run(2, 0, la, half, npair)
        assert float(torch.linalg.norm(acc - full) / torch.linalg.norm(full)) < TOL()


def test_fusion_requires_absorbable_one_body_term():
    """a non-symmetric or wrong-class h1 cannot be folded into the compressed operand:
    the three-kernel path is used and the result is still exact"""
    from fqe_b200 import synth
    from fqe_b200.fqe_data import DenseOperator, FqeData
    na, nb, norb = 3, 3, 6
    h1, h2 = synth.integrals(norb, "real8")
    rng = np.random.default_rng(5)
    h1n = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
    g = O.graph(na, nb, norb)
    c = synth.state(g.lena, g.lenb, seed=4)
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    out = d.apply((h1n, h2)).to_numpy()
    assert O.rel_err(out, O.sigma_restricted(g, c, h1n, h2)) < TOL()
    out = d.apply((h1n.real.copy(), h2)).to_numpy()     # real but non-symmetric h1
    assert O.rel_err(out, O.sigma_restricted(g, c, h1n.real, h2)) < TOL()


def test_one_body_shards():
    """one-body operator: full range (compact-list kernel) equals the sum of row shards and of
    pair slices (gather kernel with stores disabled)"""
    from fqe_b200 import lib as L
    from fqe_b200 import synth
    from fqe_b200.fqe_data import DenseOperator, FqeData
    na, nb, norb = 4, 3, 7
    h1, _ = synth.integrals(norb, "herm")
    g = O.graph(na, nb, norb)
    c = synth.state(g.lena, g.lenb, seed=21)
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    op = DenseOperator(norb, h1, None)
    ref = O.sigma_one_body(g, c, h1)
    full = d.apply_operator(op)
    assert O.rel_err(full.cpu().numpy(), ref) < TOL()
    la, npair = d.lena(), op.npair
    rows = d.apply_operator(op, row_range=(0, 11)) + d.apply_operator(op, row_range=(11, la))
    assert O.rel_err(rows.cpu().numpy(), ref) < TOL()
    pairs = d.apply_operator(op, pair_range=(0, 20)) + d.apply_operator(op, pair_range=(20, npair))
    assert O.rel_err(pairs.cpu().numpy(), ref) < TOL()


def test_host_apply_stream_pipelines_independent_builds():
    """HostApplyStream: five builds on different inputs through two device slots; every
    result equals the unpipelined apply of its own input"""
    from fqe_b200 import synth
    from fqe_b200.distributed import HostApplyStream
    from fqe_b200.fqe_data import DenseOperator, FqeData
    na, nb, norb = 4, 4, 8
    h1, h2 = synth.integrals(norb, "real8")
    op = DenseOperator(norb, h1, h2)
    g = O.graph(na, nb, norb)
    d = FqeData(na, nb, norb)
    ins = [torch.from_numpy(synth.state(g.lena, g.lenb, seed=100 + k)).pin_memory()
           for k in range(5)]
    outs = [torch.empty((g.lena, g.lenb), dtype=torch.complex128).pin_memory() for _ in range(5)]
    pipe = HostApplyStream(d)
    for c, s in zip(ins, outs):
        pipe.submit(op, c, s)
    pipe.drain()
    for c, s in zip(ins, outs):
        d.set_wfn(strategy="from_data", raw_data=c.numpy())
        ref = d.apply_operator(op).cpu().numpy()
        assert O.rel_err(s.numpy(), ref) < 1e-14


@pytest.mark.parametrize("cfg", [(0, 0, 3), (1, 0, 1), (0, 1, 1), (2, 2, 2), (0, 0, 1)])
def test_degenerate_sectors(cfg):
    """vacuum, single orbital, completely filled: one-determinant sectors through every
    entry point (no electrons: the absorbed one-body operand cannot be used)"""
    import fqe_b200
    from fqe_b200 import synth
    from fqe_b200.fqe_data import FqeData
    na, nb, norb = cfg
    h1, h2 = synth.integrals(norb, "herm", seed=1)
    g = O.graph(na, nb, norb)
    c = synth.state(g.lena, g.lenb, seed=2)
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    ref = O.sigma_restricted(g, c, h1, h2)
    out = d.apply((h1, h2)).to_numpy()
    assert np.abs(out - ref).max() < 1e-13
    out = d.apply((h1,)).to_numpy()
    assert np.abs(out - O.sigma_one_body(g, c, h1)).max() < 1e-13
    wfn = fqe_b200.Wavefunction([[na + nb, na - nb, norb]])
    wfn.set_wfn(strategy="from_data", raw_data={(na + nb, na - nb): c})
    ham = fqe_b200.get_restricted_hamiltonian((h1, h2), e_0=0.5)
    ev = wfn.time_evolve(0.05, ham).get_coeff((na + nb, na - nb))
    refev, _ = O.time_evolve_restricted(g, c, 0.05, h1, h2, 0.5)
    assert np.abs(ev - refev).max() < 1e-13
    quad = wfn.time_evolve(0.3, fqe_b200.get_restricted_hamiltonian((h1,)))
    assert np.abs(quad.get_coeff((na + nb, na - nb)) -
                  O.time_evolve_quadratic(g, c, 0.3, h1)).max() < 1e-13
    r1, r2 = d.rdm12()
    o1, o2 = O.rdm12(g, c)
    assert np.abs(r1 - o1).max() < 1e-13 and np.abs(r2 - o2).max() < 1e-13


@pytest.mark.parametrize("kind", ["real8", "herm"])
def test_deferred_scatter_by_target_slices(kind):
    """fqeb_sigma_restricted_deferred + fqeb_scatter_rows (the multi-GPU overlap path): the last
    chunk's scatter issued in slices of target rows gives bitwise the same sigma as the plain
    build, for chunked builds, row shards and pair shards; a one-body operator defers nothing"""
    from fqe_b200 import lib as L
    from fqe_b200.fqe_data import DenseOperator, FqeData
    na, nb, norb = 4, 4, 8
    g, c, h1, h2 = _case(na, nb, norb, kind)
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    op = DenseOperator(norb, h1, h2)
    lib = L.load()
    la, npair = d.lena(), op.npair

    def both(rows_per_chunk, r0, r1, p0, p1, cuts):
        nbytes = int(lib.fqeb_sigma_workspace_bytes(d._core.handle, op.handle, rows_per_chunk,
                                                    p0, p1))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device="cuda")
        plain = torch.empty_like(d.coeff)
        L.call("fqeb_sigma_restricted", d._core.handle, op.handle, d.coeff.data_ptr(),
               plain.data_ptr(), ws.data_ptr(), nbytes, r0, r1, p0, p1, None)
        torch.cuda.synchronize()
        out = torch.empty_like(d.coeff)
        pend = L.PendingScatter()
        L.call("fqeb_sigma_restricted_deferred", d._core.handle, op.handle, d.coeff.data_ptr(),
               out.data_ptr(), ws.data_ptr(), nbytes, r0, r1, p0, p1, ctypes.byref(pend), None)
        for x0, x1 in zip(cuts[:-1], cuts[1:]):
            L.call("fqeb_scatter_rows", d._core.handle, ctypes.byref(pend), x0, x1,
                   out.data_ptr(), None)
        torch.cuda.synchronize()
        return plain, out, pend

    for rows, r0, r1, p0, p1 in [(la, 0, la, 0, npair), (9, 0, la, 0, npair), (5, 11, 47, 0, npair),
                                 (la, 0, la, 0, (npair // 2) & ~1), (4, 3, 3, 0, npair)]:
        plain, out, pend = both(rows, r0, r1, p0, p1, [0, 1, 20, 21, la])
        assert torch.equal(plain, out), (rows, r0, r1, p0, p1)
        assert (pend.nrows > 0) == (r1 > r0 and p1 > p0)
    # through the Python layer
    sig, pend = d.apply_operator(op, defer_last_scatter=True)
    for x0, x1 in [(0, 30), (30, la)]:
        d.finish_scatter(pend, x0, x1, sig)
    assert torch.equal(sig, d.apply_operator(op))
    op1 = DenseOperator(norb, h1, None)
    sig, pend = d.apply_operator(op1, defer_last_scatter=True)
    assert pend.nrows == 0
    d.finish_scatter(pend, 0, la, sig)
    assert torch.equal(sig, d.apply_operator(op1))
    # bad slices are errors
    sig, pend = d.apply_operator(op, defer_last_scatter=True)
    assert lib.fqeb_scatter_rows(d._core.handle, ctypes.byref(pend), 5, 3, sig.data_ptr(), None) \
        == L.ERR_INVALID


def test_host_buffer_apply_single_gpu():
    """sharded_apply_host on one GPU (pinned host buffers in, pinned host buffers out, the
    download overlapped with the sliced scatter of the last chunk) equals the resident build"""
    from fqe_b200.distributed import ExchangeBuffers, sharded_apply_host
    from fqe_b200.fqe_data import DenseOperator, FqeData
    na, nb, norb = 4, 4, 8
    g, c, h1, h2 = _case(na, nb, norb, "real8")
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=c)
    op = DenseOperator(norb, h1, h2)
    ref = d.apply_operator(op).cpu()
    host_c = torch.from_numpy(np.ascontiguousarray(c)).pin_memory()
    host_s = torch.empty_like(host_c).pin_memory()
    bufs = ExchangeBuffers(d, 1, 0)
    for _ in range(2):   # buffers are reusable
        host_s.zero_()
        sharded_apply_host(d, op, host_c, host_s, "det", bufs)
        assert torch.equal(host_s, ref)
