"""Pin the CPU oracle (oracle/fqe_oracle.py, oracle/ref_harness.py) against the
reference: its shipped golden vectors, its literal known-answer tables and
outputs of its public API recorded by tests/golden/make_golden.py.

Tolerance: 1e-12 relative 2-norm for floating point (the path's target is 1e-10),
exact equality for integer tables."""
import os

import numpy as np
import pytest

from oracle import fqe_oracle as O
from oracle import ref_harness as R

TOL = 1e-12
HAVE_REF = R.available()
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def shipped(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_unittest_fqe_data.npz"))


@pytest.fixture(scope="module")
def graphs(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_graphs.npz"))


@pytest.fixture(scope="module")
def api(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_api.npz"))


def _impls():
    out = [("numpy", O, O.graph)]
    if HAVE_REF:
        out.append(("ref_c", R, R.graph))
    return out


# ---- literal known answers from the reference's tests/fci_graph_test.py ------
def test_string_order_is_z_address_not_integer():
    # (N=6, n=3) table starts 7, 11, 19, 35, 13, 21
    assert O.build_strings(3, 6)[:6].tolist() == [7, 11, 19, 35, 13, 21]


def test_string_address_known_answer():
    # tests/fci_graph_test.py:121  _build_string_address(4, 8, [1,2,3,7]) == 38
    s = np.array([(1 << 1) | (1 << 2) | (1 << 3) | (1 << 7)], dtype=np.uint64)
    assert int(O.string_addresses(s, 8, 4)[0]) == 38


def test_maps_norb4_known_answer():
    # tests/fci_graph_test.py:145-207 (2 alpha, 1 beta, 4 orbitals): spot values
    g = O.graph(2, 1, 4)
    assert g.astr.tolist() == [3, 5, 9, 6, 10, 12]
    assert g.bstr.tolist() == [1, 2, 4, 8]
    assert g.alpha_map[(0, 0)].tolist() == [[0, 0, 1], [1, 1, 1], [2, 2, 1]]
    assert g.alpha_map[(0, 1)].tolist() == [[3, 1, 1], [4, 2, 1]]
    assert g.alpha_map[(0, 2)].tolist() == [[3, 0, -1], [5, 2, 1]]
    assert g.beta_map[(3, 0)].tolist() == [[0, 3, 1]]


@pytest.mark.parametrize("cfg", [(2, 1, 4), (2, 3, 6), (4, 4, 8), (3, 5, 8),
                                 (0, 2, 5), (5, 5, 10), (1, 1, 1), (3, 3, 3),
                                 (2, 2, 7)])
def test_graph_tables_bit_exact(graphs, cfg):
    na, nb, norb = cfg
    k = f"{na}_{nb}_{norb}"
    for name, _, mk in _impls():
        g = mk(na, nb, norb)
        assert np.array_equal(g.astr, graphs[k + "_astr"]), name
        assert np.array_equal(g.bstr, graphs[k + "_bstr"]), name
        assert np.array_equal(g.dexca, graphs[k + "_dexca"]), name
        assert np.array_equal(g.dexcb, graphs[k + "_dexcb"]), name
        for i in range(norb):
            for j in range(norb):
                assert np.array_equal(g.alpha_map[(i, j)],
                                      graphs[f"{k}_amap_{i}_{j}"]), (name, i, j)
                assert np.array_equal(g.beta_map[(i, j)],
                                      graphs[f"{k}_bmap_{i}_{j}"]), (name, i, j)


# ---- the reference's shipped goldens (tests/fqe_data_test.py:268-280,443-456,522-565)
@pytest.mark.parametrize("cfg", [(2, 3, 6), (2, 1, 4), (1, 1, 2)])
def test_shipped_sigma_goldens(shipped, cfg):
    na, nb, norb = cfg
    s = f"{na:02d}{nb:02d}{norb:02d}"
    for name, mod, mk in _impls():
        g = mk(na, nb, norb)
        shp = (g.lena, g.lenb)
        c = (shipped["cr" + s] + 1j * shipped["ci" + s]).reshape(shp)
        h1 = shipped["h1" + s].reshape((norb,) * 2)
        h2 = shipped["h2" + s].reshape((norb,) * 4)

        def ref(tag):
            return (shipped[f"cr{s}_{tag}"] + 1j * shipped[f"ci{s}_{tag}"]).reshape(shp)

        assert O.rel_err(mod.sigma_restricted(g, c, h1, h2), ref("12")) < TOL, name
        assert O.rel_err(mod.sigma_restricted(g, c, np.zeros_like(h1), h2),
                         ref("2")) < TOL, name
        assert O.rel_err(mod.sigma_restricted(g, c, h1, np.zeros_like(h2)),
                         ref("1")) < TOL, name
    assert O.rel_err(O.sigma_one_body(O.graph(na, nb, norb), c, h1), ref("1")) < TOL


@pytest.mark.parametrize("cfg", [(2, 3, 6), (2, 1, 4)])
def test_shipped_dc_evolve_golden(shipped, cfg):
    na, nb, norb = cfg
    s = f"{na:02d}{nb:02d}{norb:02d}"
    for name, mod, mk in _impls():
        g = mk(na, nb, norb)
        shp = (g.lena, g.lenb)
        c = (shipped["cr" + s] + 1j * shipped["ci" + s]).reshape(shp)
        dmat = shipped["dmat" + s].reshape(norb, norb)
        ref = (shipped[f"cr{s}_dc"] + 1j * shipped[f"ci{s}_dc"]).reshape(shp)
        out = mod.dc_evolve(g, c, np.zeros(norb), -0.1j * dmat)
        assert O.rel_err(out, ref) < TOL, name


# ---- outputs of the reference public API -----------------------------------------
@pytest.mark.parametrize("tag", list("abcde"))
def test_api_goldens(api, tag):
    n, sz, norb = [int(x) for x in api[f"{tag}_meta"]]
    na, nb = (n + sz) // 2, (n - sz) // 2
    e0 = complex(api[f"{tag}_e0"][0])
    t = float(api[f"{tag}_t"][0])
    h1, h2, c0 = api[f"{tag}_h1"], api[f"{tag}_h2"], api[f"{tag}_c0"]
    vij = api[f"{tag}_vij"]
    for name, mod, mk in _impls():
        g = mk(na, nb, norb)
        assert O.rel_err(mod.sigma_restricted(g, c0, h1, h2), api[f"{tag}_sigma"]) < TOL, name
        assert O.rel_err(mod.dc_apply(g, c0, np.zeros(norb), vij) + e0 * c0,
                         api[f"{tag}_dc_apply"]) < TOL, name
        ev = mod.dc_evolve(g, c0, np.zeros(norb), -0.1j * vij) * np.exp(-0.1j * e0)
        assert O.rel_err(ev, api[f"{tag}_dc_evolve"]) < TOL, name
    g = O.graph(na, nb, norb)
    assert O.rel_err(O.wfn_apply_restricted(g, c0, h1, h2, e0), api[f"{tag}_apply"]) < TOL
    d4, v4 = O.dc_tensors(-np.einsum("ij,ik,jl->ijkl", vij, np.eye(norb), np.eye(norb)))
    assert O.rel_err(O.wfn_apply_dc(g, c0, d4, v4), api[f"{tag}_dc4_apply"]) < TOL
    assert O.rel_err(O.time_evolve_dc(g, c0, 0.1, np.zeros(norb), vij, e0),
                     api[f"{tag}_dc_evolve"]) < TOL
    if f"{tag}_evolve" in api:
        out, _ = O.time_evolve_restricted(g, c0, t, h1, h2, e0)
        assert O.rel_err(out, api[f"{tag}_evolve"]) < TOL
        out, _ = O.taylor(g, c0, t, h1, h2, e0)
        assert O.rel_err(out, api[f"{tag}_taylor"]) < TOL
    if f"{tag}_cheb" in api:
        out, _ = O.chebyshev(g, c0, t, h1, h2, list(api[f"{tag}_speclim"]), e0)
        assert O.rel_err(out, api[f"{tag}_cheb"]) < 1e-11
        # independent: exact exponential; taylor alone is exact, time_evolve
        # carries the e_0 phase twice
        assert O.rel_err(api[f"{tag}_taylor"], api[f"{tag}_exact"]) < 1e-10
        assert O.rel_err(api[f"{tag}_evolve"],
                         api[f"{tag}_exact"] * np.exp(-1j * t * e0)) < 1e-10


@needs_ref
@pytest.mark.parametrize("cfg", [(3, 3, 6), (4, 3, 7), (4, 4, 8)])
def test_numpy_vs_ref_c_random(cfg):
    na, nb, norb = cfg
    rng = np.random.default_rng(77 + norb)
    g, rg = O.graph(na, nb, norb), R.graph(na, nb, norb)
    c = rng.standard_normal((g.lena, g.lenb)) + 1j * rng.standard_normal((g.lena, g.lenb))
    h1 = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
    h2 = rng.standard_normal((norb,) * 4) + 1j * rng.standard_normal((norb,) * 4)
    assert O.rel_err(O.sigma_restricted(g, c, h1, h2), R.sigma_restricted(rg, c, h1, h2)) < TOL
    d = O.dvec_spatial(g, c)
    assert np.array_equal(d, R.dvec_spatial(rg, c))
    assert O.rel_err(O.coeff_from_dvec(g, d), R.coeff_from_dvec(rg, d)) < TOL
    diag = rng.standard_normal(norb) + 1j * rng.standard_normal(norb)
    v = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
    assert O.rel_err(O.dc_apply(g, c, diag, v), R.dc_apply(rg, c, diag, v)) < TOL
    assert O.rel_err(O.dc_evolve(g, c, 0.1 * diag, 0.1 * v), R.dc_evolve(rg, c, 0.1 * diag, 0.1 * v)) < TOL


@pytest.mark.parametrize("tag", ["t3a", "t3b", "t3c"])
def test_three_body_api_goldens(api, tag):
    """dense 3-body apply restatement vs outputs of the reference's wfn.apply((h1,h2,h3))"""
    n, sz, norb = [int(x) for x in api[f"{tag}_meta"]]
    g = O.graph((n + sz) // 2, (n - sz) // 2, norb)
    out = O.sigma_restricted_123(g, api[f"{tag}_c0"], api[f"{tag}_h1"], api[f"{tag}_h2"],
                                 api[f"{tag}_h3"])
    assert O.rel_err(out, api[f"{tag}_sigma"]) < 1e-12


# ---- orbital rotation (transform) and quadratic evolution -------------------
@pytest.fixture(scope="module")
def rot(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_transform.npz"))


def _sector(z, tag):
    n, sz, norb = [int(x) for x in z[f"{tag}_meta"]]
    na = (n + sz) // 2
    return O.graph(na, n - na, norb), z[f"{tag}_c0"], float(z[f"{tag}_t"][0]), complex(
        z[f"{tag}_e0"][0])


@pytest.mark.parametrize("tag", ["ta", "tb", "tc", "td", "te", "tf"])
def test_transform_goldens(rot, tag):
    """Wavefunction.transform / quadratic time_evolve / Diagonal apply + evolve of the
    reference's public API (recorded by make_golden.py --only transform)"""
    g, c0, t, e0 = _sector(rot, tag)
    perm, low, upp, c = O.transform(g, c0, rot[f"{tag}_rot"])
    assert np.array_equal(perm, rot[f"{tag}_perm"])
    assert np.allclose(low, rot[f"{tag}_low"], atol=1e-14, rtol=0)
    assert np.allclose(upp, rot[f"{tag}_upp"], atol=1e-14, rtol=0)
    assert O.rel_err(c, rot[f"{tag}_transformed"]) < TOL
    h1, diag = rot[f"{tag}_h1"], rot[f"{tag}_diag"]
    q = O.time_evolve_quadratic(g, c0, t, h1, e0)
    assert O.rel_err(q, rot[f"{tag}_quad_evolve"]) < TOL
    assert O.rel_err(O.evolve_diagonal(g, c0, -1j * t * diag), rot[f"{tag}_evolve_diagonal"]) < TOL
    assert O.rel_err(O.apply_diagonal(g, c0, diag) + e0 * c0, rot[f"{tag}_diag_apply"]) < TOL
    assert O.rel_err(O.evolve_diagonal(g, c0, -1j * t * diag) * np.exp(-1j * t * e0),
                     rot[f"{tag}_diag_evolve"]) < TOL
    # independent of the snapshot: exact exponential of the dense quadratic Hamiltonian
    if g.lena * g.lenb <= 400:
        hm = O.dense_hamiltonian(g, h1, np.zeros((g.norb,) * 4, dtype=np.complex128), e0)
        w, v = np.linalg.eigh(hm)
        exact = (v * np.exp(-1j * t * w)) @ (v.conj().T @ c0.reshape(-1))
        assert O.rel_err(q.reshape(-1), exact) < 1e-11


def test_transform_round_trip():
    """transform(U) realises the rotation U P (P = LU pivot permutation); rotating back with
    (U P)^H and the same L, U factors, as time_evolve does, is the identity; norms are kept"""
    from scipy.linalg import expm
    g = O.graph(2, 1, 4)
    rng = np.random.default_rng(3)
    c0 = rng.standard_normal((g.lena, g.lenb)) + 1j * rng.standard_normal((g.lena, g.lenb))
    assert O.rel_err(O.transform(g, c0, np.identity(4, dtype=np.complex128))[3], c0) < TOL
    a = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    u = expm(-0.5j * (a + a.conj().T))
    perm, low, upp, c1 = O.transform(g, c0, u)
    assert abs(np.linalg.norm(c1) - np.linalg.norm(c0)) < 1e-12
    assert O.rel_err(c1, c0) > 1e-2
    back = O.transform(g, c1, (u @ perm).T.conj(), low, upp)[3]
    assert O.rel_err(back, c0) < 1e-11


# ---- individual n-body operators --------------------------------------------
@pytest.fixture(scope="module")
def nbody(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_nbody.npz"))


@pytest.mark.parametrize("tag", ["na", "nb", "nc", "nd"])
def test_individual_nbody_goldens(nbody, tag):
    """FqeData.apply_individual_nbody / evolve_individual_nbody_nontrivial /
    evolve_inplace_individual_nbody_trivial of the reference, and an independent brute-force
    application of the same ladder-operator product"""
    import ast
    ops = [ast.literal_eval(str(o)) for o in nbody["ops"]]
    zc, time = complex(nbody["coeff"][0]), float(nbody["time"][0])
    n, sz, norb = [int(x) for x in nbody[f"{tag}_meta"]]
    na = (n + sz) // 2
    g, c0 = O.graph(na, n - na, norb), nbody[f"{tag}_c0"]
    seen = 0
    for k, (da, ua, db, ub) in enumerate(ops):
        if f"{tag}_apply{k}" not in nbody:
            continue
        seen += 1
        ref = nbody[f"{tag}_apply{k}"]
        assert np.abs(O.apply_individual_nbody(g, c0, zc, da, ua, db, ub) - ref).max() < 1e-15
        seq = [(2 * o, 1) for o in da] + [(2 * o, 0) for o in ua] + \
            [(2 * o + 1, 1) for o in db] + [(2 * o + 1, 0) for o in ub]
        assert np.abs(O.ladder_sequence_apply(g, c0, seq, zc) - ref).max() < 1e-15
        if f"{tag}_trivial{k}" in nbody:
            out = O.evolve_individual_trivial(g, c0, time, zc, da, db)
            assert O.rel_err(out, nbody[f"{tag}_trivial{k}"]) < TOL
        else:
            out = O.evolve_individual_nontrivial(g, c0, time, zc, da, ua, db, ub)
            assert O.rel_err(out, nbody[f"{tag}_evolve{k}"]) < TOL
    assert seen >= 6
    if f"{tag}_w_pair_apply" in nbody:
        t_op = (zc, [(2, 1), (0, 0)], [(1, 1), (3, 0)])
        t_dag = (np.conj(zc), [(0, 1), (2, 0)], [(3, 1), (1, 0)])
        assert O.rel_err(O.sparse_apply(g, c0, [t_op, t_dag], 0.25),
                         nbody[f"{tag}_w_pair_apply"]) < TOL
        ev = O.evolve_individual_nontrivial(g, c0, time, zc, [2], [0], [1], [3])
        assert O.rel_err(ev * np.exp(-0.25j * time), nbody[f"{tag}_w_pair_evolve"]) < TOL


# ---- RDMs ------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["ra", "rb", "rc", "rd", "re"])
def test_rdm_goldens(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, "ref_rdm.npz"))
    n, sz, norb = [int(x) for x in z[f"{tag}_meta"]]
    na = (n + sz) // 2
    g = O.graph(na, n - na, norb)
    r1, r2 = O.rdm12(g, z[f"{tag}_ket"])
    t1, t2 = O.rdm12(g, z[f"{tag}_ket"], z[f"{tag}_bra"])
    for got, key in ((r1, "rdm1"), (r1, "rdm12_1"), (r2, "rdm12_2"), (t1, "trdm1"),
                     (t1, "trdm12_1"), (t2, "trdm12_2")):
        assert O.rel_err(got, z[f"{tag}_{key}"]) < TOL


def test_spin_block_rotation_golden(golden_dir):
    """Wavefunction.transform with a block-diagonal 2norb x 2norb rotation (different alpha and
    beta unitaries), recorded from the reference (wavefunction.py:929-957)"""
    z = np.load(os.path.join(golden_dir, "ref_transform.npz"))
    n, sz, norb = [int(x) for x in z["tz_meta"]]
    na = (n + sz) // 2
    g = O.graph(na, n - na, norb)
    rot = z["tz_rot"]
    fa, fb = O.lu_factors(rot[:norb, :norb]), O.lu_factors(rot[norb:, norb:])
    out = O.apply_columns_recursive(g, z["tz_c0"], O.column_operator(fa[3], fa[4]),
                                    O.column_operator(fb[3], fb[4]))
    assert O.rel_err(out, z["tz_transformed"]) < TOL
    assert np.array_equal(z["tz_perm"][:norb, :norb], fa[0])
    assert np.array_equal(z["tz_perm"][norb:, norb:], fb[0])
    assert np.allclose(z["tz_low"][norb:, norb:], fb[1], atol=1e-14, rtol=0)
    assert np.allclose(z["tz_upp"][:norb, :norb], fa[2], atol=1e-14, rtol=0)


def test_oracle_against_reference_at_norb10(golden_dir):
    """The numpy restatement against the reference's own outputs at a size beyond the small
    fixtures (tests/golden/ref_large.npz, norb = 10, half filling: 63 504 determinants): rdm12,
    plain and transition."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(golden_dir), "..", "openfermion-fqe_b200"))
    from fqe_b200 import synth
    z = np.load(os.path.join(golden_dir, "ref_large.npz"))
    if "rdm10_meta" not in z.files:
        pytest.skip("rdm10 not in ref_large.npz")
    n, sz, norb = [int(x) for x in z["rdm10_meta"]]
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    ket = synth.state(la, lb, seed=synth.seed_for(norb, 54))
    bra = synth.state(la, lb, seed=synth.seed_for(norb, 55))
    g = O.graph(na, nb, norb)
    for tag, (r1, r2) in (("rdm10", O.rdm12(g, ket)), ("trdm10", O.rdm12(g, ket, bra))):
        assert O.rel_err(r1, z[f"{tag}_1"]) < 1e-12, tag
        flat = np.asarray(r2).reshape(-1)
        assert O.rel_err(flat[z[f"{tag}_2_idx"]], z[f"{tag}_2_val"]) < 1e-12, tag
        assert abs(np.linalg.norm(r2) - z[f"{tag}_2_norm"][0]) < 1e-11 * z[f"{tag}_2_norm"][0]
