#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It (1) builds the UNMODIFIED reference (libfqe.so + the Cython shim) from a
scratch copy of /root/reference/src under /tmp, imports it behind permissive
``openfermion``/``cirq`` stand-ins (neither is installed here; the dense hot
path never touches them), (2) re-exports the reference's own golden vectors for
this path (tests/unittest_data/fqe_data/*.npy) and (3) records outputs of the
reference's public API (fqe.Wavefunction.apply / time_evolve /
apply_generated_unitary, FqeData.apply, FciGraph tables) on seeded inputs.

Outputs (all small, committed):
    ref_unittest_fqe_data.npz   reference's shipped goldens, verbatim values
    ref_graphs.npz              strings / maps / dexc from reference FciGraph
    ref_api.npz                 Wavefunction-level inputs + outputs
    ref_hring12.npz             BASELINE config 0 (profiling/profile_H_ring.py): integrals read
                                from profiling/Hring_12.hdf5 with tests/golden/hdf5_mini.py,
                                energies, and signatures (norms, 512 sampled coefficients, 8
                                random projections) of H|HF> and of the state evolved to t=0.1
    ref_transform.npz           SURVEY 8f rank 1: Wavefunction.transform (LU column rotations),
                                quadratic time_evolve, Diagonal apply / evolve, evolve_diagonal
                                (`--only transform` regenerates just this file)
    ref_nbody.npz               SURVEY 8f rank 2: individual n-body apply / exact evolution through
                                FqeData and through Wavefunction + SparseHamiltonian (`--only nbody`)
    ref_rdm.npz                 SURVEY 8f rank 4: FqeData.rdm1 / rdm12, plain and transition (`--only rdm`)
    ref_wick.npz                Wavefunction.rdm / expectationValue with operator strings, every
                                spin-free rank-1 / rank-2 ordering (`--only wick`)
    ref_wfn_save.bin / .npz     a file written by the reference's Wavefunction.save and its
                                coefficients (`--only wfnio`)
"""
import os
import shutil
import subprocess
import sys
import types

import numpy as np

REF = "/root/reference"
SCRATCH = "/tmp/fqe_ref_build"
HERE = os.path.dirname(os.path.abspath(__file__))


def build_reference():
    src = os.path.join(SCRATCH, "src")
    so = os.path.join(src, "fqe", "lib", "libfqe.so")
    if not os.path.exists(so):
        if os.path.exists(SCRATCH):
            shutil.rmtree(SCRATCH)
        shutil.copytree(os.path.join(REF, "src"), src)
        libdir = os.path.join(src, "fqe", "lib")
        cfiles = ["macros.c", "mylapack.c", "fci_graph.c", "fqe_data.c",
                  "cirq_utils.c", "wick.c", "bitstring.c", "binom.c"]
        subprocess.check_call(
            ["/usr/bin/gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-I."] +
            cfiles + ["-o", "libfqe.so", "-lm"], cwd=libdir)
    ext = [f for f in os.listdir(os.path.join(src, "fqe", "lib"))
           if f.startswith("fqe_data.") and f.endswith(".so")]
    if not ext:
        setup = os.path.join(SCRATCH, "build_ext.py")
        with open(setup, "w") as fh:
            fh.write(
                "from setuptools import setup, Extension\n"
                "from Cython.Build import cythonize\n"
                "ext = Extension('fqe.lib.fqe_data', ['src/fqe/lib/_fqe_data.pyx'], language='c')\n"
                "setup(name='fqe_ref', package_dir={'': 'src'},\n"
                "      ext_modules=cythonize([ext], compiler_directives={'language_level': '3'}),\n"
                "      script_args=['build_ext', '--inplace'])\n")
        env = dict(os.environ, CC="/usr/bin/gcc")
        subprocess.check_call([sys.executable, setup], cwd=SCRATCH, env=env)
    return src


def install_stubs():
    class _Dummy:
        def __init__(self, *a, **k):
            pass

    def mk(name):
        m = types.ModuleType(name)
        m.__path__ = []
        m.__getattr__ = lambda attr: type(attr, (_Dummy,), {})
        return m

    for n in ["openfermion", "openfermion.ops", "openfermion.utils",
              "openfermion.transforms", "openfermion.transforms.opconversions",
              "openfermion.chem", "openfermion.chem.molecular_data",
              "openfermion.linalg", "openfermion.circuits",
              "openfermion.circuits.primitives", "cirq", "cirq.ops",
              "cirq.ops.pauli_string"]:
        sys.modules[n] = mk(n)
    sys.modules["openfermion"].up_index = lambda i: 2 * i
    sys.modules["openfermion"].down_index = lambda i: 2 * i + 1


def synth(norb, seed, kind):
    """Seeded synthetic integrals (SURVEY 8d recipe)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((norb, norb))
    if kind == "real8":
        h1 = 0.5 * (a + a.T)
        v = rng.standard_normal((norb,) * 4)
        v = v + v.transpose(1, 0, 2, 3)
        v = v + v.transpose(0, 1, 3, 2)
        v = v + v.transpose(2, 3, 0, 1)
        h2 = np.einsum("ijkl->ikjl", v) * 0.1
        return h1.astype(np.complex128), h2.astype(np.complex128)
    if kind == "herm":
        b = rng.standard_normal((norb, norb))
        h1 = 0.5 * ((a + 1j * b) + (a + 1j * b).conj().T)
        w = rng.standard_normal((norb,) * 4) + 1j * rng.standard_normal((norb,) * 4)
        h2 = 0.05 * (w + w.conj().transpose(3, 2, 1, 0))
        return h1, h2
    if kind == "general":  # no symmetry at all, complex
        b = rng.standard_normal((norb, norb))
        w = rng.standard_normal((norb,) * 4) + 1j * rng.standard_normal((norb,) * 4)
        return (a + 1j * b), 0.05 * w
    raise ValueError(kind)


def rand_state(shape, rng):
    c = rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)
    return c / np.linalg.norm(c)


def transform_goldens(fqe):
    """Orbital rotations and quadratic evolution through the reference's public API."""
    import copy
    from scipy.linalg import expm
    out = {}
    cases = [("ta", 4, 0, 4, False), ("tb", 5, 1, 6, False), ("tc", 3, -1, 5, True),
             ("td", 6, 0, 6, False), ("te", 2, 2, 5, False), ("tf", 8, 0, 8, False)]
    for tag, n, sz, norb, real in cases:
        rng = np.random.default_rng(20260600 + 100 * norb + ord(tag[1]))
        a = rng.standard_normal((norb, norb))
        b = np.zeros((norb, norb)) if real else rng.standard_normal((norb, norb))
        if real:   # real orthogonal rotation
            rot = expm(0.4 * (a - a.T))
        else:
            k = 0.5 * ((a + 1j * b) + (a + 1j * b).conj().T)
            rot = expm(-0.7j * k)
        h1 = 0.5 * ((a + 1j * b) + (a + 1j * b).conj().T)
        if real:
            h1 = h1.real.astype(np.complex128)
        diag = rng.standard_normal(norb)
        wfn = fqe.Wavefunction([[n, sz, norb]])
        shape = wfn.get_coeff((n, sz)).shape
        c0 = rand_state(shape, rng)
        wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0.copy()})
        t, e0 = 0.37, -0.4
        out[f"{tag}_meta"] = np.array([n, sz, norb], dtype=np.int64)
        out[f"{tag}_c0"], out[f"{tag}_rot"], out[f"{tag}_h1"], out[f"{tag}_diag"] = c0, rot, h1, diag
        out[f"{tag}_t"], out[f"{tag}_e0"] = np.array([t]), np.array([e0])
        perm, low, upp, res = copy.deepcopy(wfn).transform(rot)
        out[f"{tag}_perm"], out[f"{tag}_low"], out[f"{tag}_upp"] = perm, low, upp
        out[f"{tag}_transformed"] = res.get_coeff((n, sz))
        ham = fqe.get_restricted_hamiltonian((h1,), e_0=e0)
        out[f"{tag}_quad_evolve"] = wfn.time_evolve(t, ham).get_coeff((n, sz))
        out[f"{tag}_quad_apply"] = wfn.apply(ham).get_coeff((n, sz))
        dham = fqe.get_diagonal_hamiltonian(diag.astype(np.complex128), e_0=e0)
        out[f"{tag}_diag_apply"] = wfn.apply(dham).get_coeff((n, sz))
        out[f"{tag}_diag_evolve"] = wfn.time_evolve(t, dham).get_coeff((n, sz))
        arr = (-1j * t * diag).astype(np.complex128)
        out[f"{tag}_evolve_diagonal"] = wfn.sector((n, sz)).evolve_diagonal(arr)
        # exact answer for the quadratic evolution from the dense one-particle picture:
        # <x| exp(-itH) |psi> checked through U = exp(-i t h1) as an orbital rotation
        out[f"{tag}_quad_unitary"] = expm(-1j * t * h1)
    # block-diagonal 2norb x 2norb rotation: different unitaries for alpha and beta
    n, sz, norb = 5, 1, 6
    rng = np.random.default_rng(20260699)

    def unitary():
        a = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
        return expm(-0.6j * (a + a.conj().T))

    big = np.zeros((2 * norb, 2 * norb), dtype=np.complex128)
    big[:norb, :norb], big[norb:, norb:] = unitary(), unitary()
    wfn = fqe.Wavefunction([[n, sz, norb]])
    c0 = rand_state(wfn.get_coeff((n, sz)).shape, rng)
    wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0.copy()})
    perm, low, upp, res = wfn.transform(big)
    out["tz_meta"] = np.array([n, sz, norb], dtype=np.int64)
    out["tz_c0"], out["tz_rot"] = c0, big
    out["tz_perm"], out["tz_low"], out["tz_upp"] = perm, low, upp
    out["tz_transformed"] = res.get_coeff((n, sz))
    np.savez_compressed(os.path.join(HERE, "ref_transform.npz"), **out)
    print("ref_transform.npz", os.path.getsize(os.path.join(HERE, "ref_transform.npz")), "bytes")


def nbody_goldens(fqe):
    """Individual n-body operators through FqeData and, with the internal
    operator lists set by hand (openfermion's FermionOperator is not available here),
    through Wavefunction.apply / time_evolve with a SparseHamiltonian."""
    from fqe.hamiltonians.sparse_hamiltonian import SparseHamiltonian
    from fqe.hamiltonians.hamiltonian import Hamiltonian

    def sparse(operators, e_0=0.0):
        sh = SparseHamiltonian.__new__(SparseHamiltonian)
        Hamiltonian.__init__(sh, e_0=e_0)
        sh._operators = [(c, list(a), list(b)) for c, a, b in operators]
        sh._conserve_spin = True
        sh._rank = max(len(a) + len(b) for _, a, b in operators)
        return sh

    ops = [([1], [0], [], []), ([], [], [2], [1]), ([2], [0], [1], [3]),
           ([3, 1], [2, 0], [], []), ([1], [1], [2], [0]), ([2, 0], [2, 1], [1], [1]),
           ([0], [0], [1], [1]), ([2, 1], [1, 0], [3], [2]), ([3], [3], [], [])]
    out = {"ops": np.array([repr(o) for o in ops])}
    coeff, time = 0.3 - 0.2j, 0.41
    out["coeff"], out["time"] = np.array([coeff]), np.array([time])
    for tag, n, sz, norb in [("na", 4, 0, 4), ("nb", 5, 1, 6), ("nc", 6, 0, 6), ("nd", 3, -1, 5)]:
        rng = np.random.default_rng(20260700 + 100 * norb + ord(tag[1]))
        wfn = fqe.Wavefunction([[n, sz, norb]])
        shape = wfn.get_coeff((n, sz)).shape
        c0 = rand_state(shape, rng)
        wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0.copy()})
        out[f"{tag}_meta"], out[f"{tag}_c0"] = np.array([n, sz, norb], dtype=np.int64), c0
        sec = wfn.sector((n, sz))
        for k, (da, ua, db, ub) in enumerate(ops):
            if max(da + ua + db + ub) >= norb:
                continue
            out[f"{tag}_apply{k}"] = sec.apply_individual_nbody(coeff, da, ua, db, ub).coeff
            if da == ua and db == ub:
                tmp = fqe.Wavefunction([[n, sz, norb]])
                tmp.set_wfn(strategy="from_data", raw_data={(n, sz): c0.copy()})
                tmp.sector((n, sz)).evolve_inplace_individual_nbody_trivial(time, coeff, da, db)
                out[f"{tag}_trivial{k}"] = tmp.get_coeff((n, sz))
            else:
                out[f"{tag}_evolve{k}"] = sec.evolve_individual_nbody_nontrivial(
                    time, coeff, da, ua, db, ub).coeff
        # Wavefunction level: T + T^+ for T = c a+_2a a_0a a+_1b a_3b, one number operator,
        # and a three-operator Hermitian sum that goes through the Taylor series
        t_op = (coeff, [(2, 1), (0, 0)], [(1, 1), (3, 0)])
        t_dag = (np.conj(coeff), [(0, 1), (2, 0)], [(3, 1), (1, 0)])
        if norb > 3:
            pair = sparse([t_op, t_dag], e_0=0.25)
            out[f"{tag}_w_pair_apply"] = wfn.apply(pair).get_coeff((n, sz))
            out[f"{tag}_w_pair_evolve"] = wfn.time_evolve(time, pair).get_coeff((n, sz))
            num = sparse([(0.7, [(1, 1), (1, 0)], [(0, 1), (0, 0)])], e_0=-0.5)
            out[f"{tag}_w_num_apply"] = wfn.apply(num).get_coeff((n, sz))
            out[f"{tag}_w_num_evolve"] = wfn.time_evolve(time, num).get_coeff((n, sz))
            three = sparse([t_op, t_dag, (0.4, [(1, 1), (1, 0)], [])], e_0=0.05)
            out[f"{tag}_w_three_apply"] = wfn.apply(three).get_coeff((n, sz))
            out[f"{tag}_w_three_evolve"] = wfn.time_evolve(0.05, three).get_coeff((n, sz))
    np.savez_compressed(os.path.join(HERE, "ref_nbody.npz"), **out)
    print("ref_nbody.npz", os.path.getsize(os.path.join(HERE, "ref_nbody.npz")), "bytes")


def rdm_goldens(fqe):
    """1- and 2-particle (transition) RDMs of single sectors, FqeData.rdm1 / rdm12
    (SURVEY 8f rank 4); covers the half-filling and the low-filling algorithm."""
    out = {}
    for tag, n, sz, norb in [("ra", 4, 0, 4), ("rb", 5, 1, 6), ("rc", 2, 0, 8), ("rd", 3, -1, 5),
                             ("re", 6, 0, 6)]:
        rng = np.random.default_rng(20260800 + 100 * norb + ord(tag[1]))
        wfn = fqe.Wavefunction([[n, sz, norb]])
        shape = wfn.get_coeff((n, sz)).shape
        ket, bra = rand_state(shape, rng), rand_state(shape, rng)
        wfn.set_wfn(strategy="from_data", raw_data={(n, sz): ket.copy()})
        bwfn = fqe.Wavefunction([[n, sz, norb]])
        bwfn.set_wfn(strategy="from_data", raw_data={(n, sz): bra.copy()})
        sec, bsec = wfn.sector((n, sz)), bwfn.sector((n, sz))
        out[f"{tag}_meta"] = np.array([n, sz, norb], dtype=np.int64)
        out[f"{tag}_ket"], out[f"{tag}_bra"] = ket, bra
        (out[f"{tag}_rdm1"],) = sec.rdm1()
        (out[f"{tag}_trdm1"],) = sec.rdm1(bsec)
        out[f"{tag}_rdm12_1"], out[f"{tag}_rdm12_2"] = sec.rdm12()
        out[f"{tag}_trdm12_1"], out[f"{tag}_trdm12_2"] = sec.rdm12(bsec)
    np.savez_compressed(os.path.join(HERE, "ref_rdm.npz"), **out)
    print("ref_rdm.npz", os.path.getsize(os.path.join(HERE, "ref_rdm.npz")), "bytes")


def wfnio_goldens(fqe):
    """A file written by the reference's Wavefunction.save (wavefunction.py:743-765) and the
    coefficients it held: fqe_b200.wfn_io must read it without the reference installed."""
    rng = np.random.default_rng(20260900)
    n, sz, norb = 3, 1, 4
    w = fqe.Wavefunction([[n, sz, norb]])
    shape = w.get_coeff((n, sz)).shape
    c = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    w.set_wfn(strategy="from_data", raw_data={(n, sz): c.copy()})
    w.save("ref_wfn_save.bin", HERE)
    np.savez(os.path.join(HERE, "ref_wfn_save.npz"), coeff=c, n=[n], sz=[sz], norb=[norb])
    print("ref_wfn_save.bin", os.path.getsize(os.path.join(HERE, "ref_wfn_save.bin")), "bytes")


def wick_goldens(fqe):
    """Wavefunction.rdm / expectationValue with operator strings (wavefunction.py:1331-1355,
    1100-1133; wick.py): every spin-free rank-1 and rank-2 ordering the reference accepts, plain
    and transition (single elements given by digits need openfermion's parser in the reference
    and are checked against these tensors instead)."""
    out = {}
    strings = ["i^ j", "i j^", "j^ i", "i^ j^ k l", "i^ j k l^", "i j^ k^ l", "i j k^ l^",
               "k^ l^ i j", "p^ q r s^"]
    out["strings"] = np.array(strings)
    for tag, n, sz, norb in [("wa", 4, 0, 4), ("wb", 3, 1, 5), ("wc", 2, 0, 3)]:
        rng = np.random.default_rng(20261000 + 100 * norb + ord(tag[1]))
        wfn = fqe.Wavefunction([[n, sz, norb]])
        shape = wfn.get_coeff((n, sz)).shape
        ket, bra = rand_state(shape, rng), rand_state(shape, rng)
        wfn.set_wfn(strategy="from_data", raw_data={(n, sz): ket.copy()})
        bwfn = fqe.Wavefunction([[n, sz, norb]])
        bwfn.set_wfn(strategy="from_data", raw_data={(n, sz): bra.copy()})
        out[f"{tag}_meta"] = np.array([n, sz, norb], dtype=np.int64)
        out[f"{tag}_ket"], out[f"{tag}_bra"] = ket, bra
        for k, st in enumerate(strings):
            out[f"{tag}_s{k}"] = np.asarray(wfn.rdm(st))
            out[f"{tag}_t{k}"] = np.asarray(wfn.rdm(st, brawfn=bwfn))
            out[f"{tag}_e{k}"] = np.asarray(wfn.expectationValue(st, brawfn=bwfn))
    np.savez_compressed(os.path.join(HERE, "ref_wick.npz"), **out)
    print("ref_wick.npz", os.path.getsize(os.path.join(HERE, "ref_wick.npz")), "bytes")


def main():
    src = build_reference()
    install_stubs()
    sys.path.insert(0, src)
    import fqe
    import fqe.settings
    from fqe.fci_graph import FciGraph
    from fqe.fqe_data import FqeData

    fqe.settings.use_accelerated_code = True
    if "--only" in sys.argv:
        {"transform": transform_goldens, "nbody": nbody_goldens, "rdm": rdm_goldens,
         "wfnio": wfnio_goldens, "wick": wick_goldens}[
            sys.argv[sys.argv.index("--only") + 1]](fqe)
        return

    # ---- (2) the reference's shipped goldens for this path ------------------
    d = os.path.join(REF, "tests", "unittest_data", "fqe_data")
    out = {}
    for (na, nb, norb) in [(2, 3, 6), (2, 1, 4), (1, 1, 2)]:
        s = f"{na:02d}{nb:02d}{norb:02d}"
        names = [f"cr{s}", f"ci{s}", f"h1{s}", f"h2{s}"]
        names += [f"c{p}{s}_{t}" for p in "ri" for t in ("1", "2", "12")]
        if (na, nb, norb) != (1, 1, 2):
            names += [f"dmat{s}", f"cr{s}_dc", f"ci{s}_dc"]
        for nm in names:
            out[nm] = np.fromfile(os.path.join(d, nm + ".npy"))
    np.savez_compressed(os.path.join(HERE, "ref_unittest_fqe_data.npz"), **out)

    # ---- graphs ------------------------------------------------------------
    gout = {}
    for (na, nb, norb) in [(2, 1, 4), (2, 3, 6), (4, 4, 8), (3, 5, 8), (0, 2, 5),
                           (5, 5, 10), (1, 1, 1), (3, 3, 3), (2, 2, 7)]:
        g = FciGraph(na, nb, norb)
        k = f"{na}_{nb}_{norb}"
        gout[k + "_astr"] = np.asarray(g._astr, dtype=np.uint64)
        gout[k + "_bstr"] = np.asarray(g._bstr, dtype=np.uint64)
        gout[k + "_dexca"] = g._dexca
        gout[k + "_dexcb"] = g._dexcb
        for (i, j), m in g._alpha_map.items():
            gout[f"{k}_amap_{i}_{j}"] = np.asarray(m, dtype=np.int32).reshape(-1, 3)
        for (i, j), m in g._beta_map.items():
            gout[f"{k}_bmap_{i}_{j}"] = np.asarray(m, dtype=np.int32).reshape(-1, 3)
    np.savez_compressed(os.path.join(HERE, "ref_graphs.npz"), **gout)

    # ---- (3) public-API outputs ----------------------------------------------
    api = {}
    cases = [  # (tag, n, sz, norb, kind, e0, t)
        ("a", 4, 0, 4, "real8", 0.0, 0.02),
        ("b", 5, 1, 6, "herm", -1.25, 0.01),
        ("c", 6, 0, 6, "general", 0.3 + 0.0j, 0.01),
        ("d", 3, -1, 5, "real8", 2.0, 0.02),
        ("e", 8, 0, 8, "real8", -0.7, 0.004),
    ]
    for tag, n, sz, norb, kind, e0, t in cases:
        seed = 20260000 + 100 * norb + ord(tag)
        rng = np.random.default_rng(seed)
        h1, h2 = synth(norb, seed, kind)
        wfn = fqe.Wavefunction([[n, sz, norb]])
        shape = wfn.get_coeff((n, sz)).shape
        c0 = rand_state(shape, rng)
        wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0.copy()})
        ham = fqe.get_restricted_hamiltonian((h1, h2), e_0=e0)
        api[f"{tag}_meta"] = np.array([n, sz, norb], dtype=np.int64)
        api[f"{tag}_e0"] = np.array([e0], dtype=np.complex128)
        api[f"{tag}_t"] = np.array([t])
        api[f"{tag}_h1"] = h1
        api[f"{tag}_h2"] = h2
        api[f"{tag}_c0"] = c0
        # sigma through FqeData.apply (C lm path), and Wavefunction.apply (+e0)
        api[f"{tag}_sigma"] = wfn.sector((n, sz)).apply((h1, h2)).coeff
        api[f"{tag}_apply"] = wfn.apply(ham).get_coeff((n, sz))
        if kind in ("real8", "herm"):
            ev = wfn.time_evolve(t, ham)
            api[f"{tag}_evolve"] = ev.get_coeff((n, sz))
            agu = wfn.apply_generated_unitary(t, "taylor", ham)
            api[f"{tag}_taylor"] = agu.get_coeff((n, sz))
            # spectral bounds from the dense matrix for Chebyshev
            if np.prod(shape) <= 400:
                dim = int(np.prod(shape))
                hm = np.zeros((dim, dim), dtype=np.complex128)
                for k in range(dim):
                    e = np.zeros(dim, dtype=np.complex128)
                    e[k] = 1.0
                    w = fqe.Wavefunction([[n, sz, norb]])
                    w.set_wfn(strategy="from_data",
                              raw_data={(n, sz): e.reshape(shape)})
                    hm[:, k] = w.apply(ham).get_coeff((n, sz)).reshape(-1)
                ev_ = np.linalg.eigvalsh(hm)
                lim = [float(ev_[0]) - 0.1, float(ev_[-1]) + 0.1]
                api[f"{tag}_speclim"] = np.array(lim)
                ch = wfn.apply_generated_unitary(t, "chebyshev", ham,
                                                 spec_lim=lim)
                api[f"{tag}_cheb"] = ch.get_coeff((n, sz))
                exact = (np.linalg.eigh(hm)[1] * np.exp(-1j * t * np.linalg.eigvalsh(hm))
                         ) @ (np.linalg.eigh(hm)[1].conj().T @ c0.reshape(-1))
                api[f"{tag}_exact"] = exact.reshape(shape)
        # diagonal Coulomb: non-symmetric v as in the reference's own golden
        vij = 8.0 * rng.uniform(0, 1, (norb, norb)) / norb
        dch = fqe.get_diagonalcoulomb_hamiltonian(vij, e_0=e0)
        api[f"{tag}_vij"] = vij
        api[f"{tag}_dc_apply"] = wfn.apply(dch).get_coeff((n, sz))
        api[f"{tag}_dc_evolve"] = wfn.time_evolve(0.1, dch).get_coeff((n, sz))
        # 4-index DiagonalCoulomb input
        h4 = np.zeros((norb,) * 4)
        for i in range(norb):
            for j in range(norb):
                h4[i, j, i, j] = -vij[i, j]
        dch4 = fqe.get_diagonalcoulomb_hamiltonian(h4)
        api[f"{tag}_dc4_apply"] = wfn.apply(dch4).get_coeff((n, sz))
    # ---- dense three-body apply (BASELINE config 5 shape: profiling/profile_3_body.py) ------
    for tag, n, sz, norb in [("t3a", 4, 0, 4), ("t3b", 5, 1, 5), ("t3c", 6, 0, 6)]:
        rng = np.random.default_rng(20260500 + norb)
        h1 = rng.standard_normal((norb,) * 2) + 1j * rng.standard_normal((norb,) * 2)
        h2 = 0.1 * (rng.standard_normal((norb,) * 4) + 1j * rng.standard_normal((norb,) * 4))
        if tag == "t3c":   # the profiling script's tensor, h1 = h2 = 0
            idx = np.indices((norb,) * 6)
            h3 = ((idx[0] + idx[3]) * (idx[1] + idx[4]) * (idx[2] + idx[5]) * 0.002).astype(
                np.complex128)
            h1 = np.zeros_like(h1)
            h2 = np.zeros_like(h2)
        else:
            h3 = 0.05 * (rng.standard_normal((norb,) * 6) + 1j * rng.standard_normal((norb,) * 6))
        wfn = fqe.Wavefunction([[n, sz, norb]])
        shape = wfn.get_coeff((n, sz)).shape
        c0 = rand_state(shape, rng)
        wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0.copy()})
        api[f"{tag}_meta"] = np.array([n, sz, norb], dtype=np.int64)
        api[f"{tag}_h1"], api[f"{tag}_h2"], api[f"{tag}_h3"], api[f"{tag}_c0"] = h1, h2, h3, c0
        api[f"{tag}_sigma"] = wfn.apply((h1, h2, h3)).get_coeff((n, sz))
    np.savez_compressed(os.path.join(HERE, "ref_api.npz"), **api)

    # ---- (4) BASELINE.json config 0: H12 ring as run by profiling/profile_H_ring.py ---------
    sys.path.insert(0, HERE)
    from hdf5_mini import MiniHDF5
    mol = MiniHDF5(os.path.join(REF, "profiling", "Hring_12.hdf5"))
    nele, norbs = int(mol.read("n_electrons")), int(mol.read("n_orbitals"))
    sz = int(mol.read("multiplicity")) - 1
    h1 = np.array(mol.read("one_body_integrals"))
    h2 = np.array(mol.read("two_body_integrals"))
    e_nuc = float(mol.read("nuclear_repulsion"))
    h2f = np.ascontiguousarray(np.einsum("ijlk", -0.5 * h2))
    wf = fqe.Wavefunction([[nele, sz, norbs]])
    wf.set_wfn(strategy="hartree-fock")
    wf.normalize()
    ham = fqe.get_restricted_hamiltonian((h1, h2f), e_0=e_nuc)
    e_init = wf.expectationValue(ham)
    sigma = wf.apply(ham).get_coeff((nele, sz))
    t_evolve = 0.1
    evolved = wf.time_evolve(t_evolve, ham)
    e_final = evolved.expectationValue(ham)
    cev = evolved.get_coeff((nele, sz))
    rng = np.random.default_rng(20261200)
    probe = rng.standard_normal((8,) + cev.shape) + 1j * rng.standard_normal((8,) + cev.shape)
    idx = rng.choice(cev.size, size=512, replace=False)
    np.savez_compressed(
        os.path.join(HERE, "ref_hring12.npz"), meta=np.array([nele, sz, norbs]), h1=h1, h2=h2f,
        e_0=np.array([e_nuc]), hf_energy=np.array([float(mol.read("hf_energy"))]),
        e_init=np.array([e_init]), e_final=np.array([e_final]), t=np.array([t_evolve]),
        probe_seed=np.array([20261200]), sample_idx=idx,
        sigma_samples=sigma.reshape(-1)[idx], sigma_norm=np.array([np.linalg.norm(sigma)]),
        sigma_probe=np.einsum("kab,ab->k", probe.conj(), sigma),
        evolved_samples=cev.reshape(-1)[idx], evolved_norm=np.array([np.linalg.norm(cev)]),
        evolved_probe=np.einsum("kab,ab->k", probe.conj(), cev))
    print("H12 ring: E_init", e_init, "E_final", e_final, "E_HF(file)", float(mol.read("hf_energy")))
    for f in ("ref_unittest_fqe_data.npz", "ref_graphs.npz", "ref_api.npz", "ref_hring12.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    transform_goldens(fqe)
    nbody_goldens(fqe)
    rdm_goldens(fqe)
    wfnio_goldens(fqe)


if __name__ == "__main__":
    main()
