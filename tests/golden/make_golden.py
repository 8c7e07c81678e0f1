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


def main():
    src = build_reference()
    install_stubs()
    sys.path.insert(0, src)
    import fqe
    import fqe.settings
    from fqe.fci_graph import FciGraph
    from fqe.fqe_data import FqeData

    fqe.settings.use_accelerated_code = True

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


if __name__ == "__main__":
    main()
