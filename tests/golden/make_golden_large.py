#!/usr/bin/env python
"""Reference signatures AT THE SIZES BASELINE.json's configs name.

Run in the build container only (needs /root/reference; minutes of CPU per case):

    python tests/golden/make_golden_large.py [--only sigma14|sigma16|taylor14|dc16|body3|cheb12|nbody12|rdm]

The UNMODIFIED reference (built from source by make_golden.build_reference, imported behind
the openfermion/cirq stand-ins) is driven through its public API -- FqeData.apply (the C
``lm`` path, fqe_data.py:685-710), Wavefunction.time_evolve (wavefunction.py:961-1054, Taylor
548-568), apply/evolve of a DiagonalCoulomb (lib/fqe_data.c:455-602), the dense 3-body apply
(fqe_data.py:1166-1216 on profiling/profile_3_body.py's tensor), the Chebyshev propagator and
Wavefunction.transform at norb=12, individual 3-body operators at norb=12 and rdm12 (plain and
transition) at norb=10 and 12 -- on the seeded inputs of
``fqe_b200.synth`` (the ones bench.py and the -m gpu tests regenerate from their seeds).

A full state at norb=16 is 2.65 GB, so what is committed per state is a SIGNATURE (kilobytes):
  norm                      2-norm
  idx, samples              4096 coefficients at seeded flat positions
  row_norms, col_norms      2-norm of every alpha row and every beta column
  probe                     16 rank-1 projections  u_k^H S v_k,  u_k, v_k seeded complex normals
Every row and column norm pins the magnitude of every block; the random projections and the
samples pin the phases.  ``signature()`` / ``signature_error()`` below are the single definition
used by this script, by tests/test_gpu_large_goldens.py and by bench.py --verify.

Output: tests/golden/ref_large.npz
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "ref_large.npz")
NSAMPLE, NPROBE = 4096, 16


def probes(la, lb, seed):
    rng = np.random.default_rng(seed)
    idx = rng.choice(la * lb, size=min(NSAMPLE, la * lb), replace=False)
    u = rng.standard_normal((NPROBE, la)) + 1j * rng.standard_normal((NPROBE, la))
    v = rng.standard_normal((NPROBE, lb)) + 1j * rng.standard_normal((NPROBE, lb))
    return idx, u, v


def signature(state, seed):
    """state: numpy complex128 [la, lb]"""
    la, lb = state.shape
    idx, u, v = probes(la, lb, seed)
    return {
        "seed": np.array([seed], dtype=np.int64),
        "norm": np.array([np.linalg.norm(state)]),
        "samples": state.reshape(-1)[idx],
        "row_norms": np.linalg.norm(state, axis=1),
        "col_norms": np.linalg.norm(state, axis=0),
        "probe": ((u.conj() @ state) * v).sum(axis=1),
    }


def signature_torch(state, seed):
    """signature() of a CUDA complex128 tensor [la, lb], evaluated on the device (test harness
    arithmetic; only kilobytes come back to the host)."""
    import torch
    la, lb = state.shape
    idx, u, v = probes(la, lb, seed)
    dev = state.device
    ut = torch.from_numpy(u).to(dev)
    vt = torch.from_numpy(v).to(dev)
    it = torch.from_numpy(idx.astype(np.int64)).to(dev)
    probe = ((ut.conj() @ state) * vt).sum(dim=1)
    return {
        "seed": np.array([seed], dtype=np.int64),
        "norm": np.array([float(torch.linalg.norm(state).item())]),
        "samples": state.reshape(-1)[it].cpu().numpy(),
        "row_norms": torch.linalg.norm(state, dim=1).cpu().numpy(),
        "col_norms": torch.linalg.norm(state, dim=0).cpu().numpy(),
        "probe": probe.cpu().numpy(),
    }


def stored_signature(store, tag):
    keys = ("seed", "norm", "samples", "row_norms", "col_norms", "probe")
    return {k: store[f"{tag}_{k}"] for k in keys}


def signature_error(sig, got):
    """Worst relative deviation between a stored signature ``sig`` (dict-like with the keys of
    signature()) and the signature ``got`` of a candidate state; every entry is normalised by
    the reference state's 2-norm (probe entries by the 2-norm of the reference probe vector)."""
    nrm = float(sig["norm"][0])
    errs = {
        "norm": abs(float(got["norm"][0]) - nrm) / nrm,
        "samples": float(np.linalg.norm(got["samples"] - sig["samples"])) / nrm,
        "row_norms": float(np.linalg.norm(got["row_norms"] - sig["row_norms"])) / nrm,
        "col_norms": float(np.linalg.norm(got["col_norms"] - sig["col_norms"])) / nrm,
        "probe": float(np.linalg.norm(got["probe"] - sig["probe"]) /
                       np.linalg.norm(sig["probe"])),
    }
    return max(errs.values()), errs


def load_existing():
    if os.path.exists(OUT):
        with np.load(OUT) as z:
            return {k: z[k] for k in z.files}
    return {}


def put(store, tag, sig, **extra):
    for k, v in sig.items():
        store[f"{tag}_{k}"] = v
    for k, v in extra.items():
        store[f"{tag}_{k}"] = np.asarray(v)


def main():
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
    import make_golden as MG
    from fqe_b200 import synth       # imports torch: must happen before the stand-ins exist
    src = MG.build_reference()
    MG.install_stubs()
    sys.path.insert(0, src)
    import fqe
    import fqe.settings
    fqe.settings.use_accelerated_code = True
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    store = load_existing()

    def wavefunction(n, sz, norb, c0):
        wfn = fqe.Wavefunction([[n, sz, norb]])
        wfn.set_wfn(strategy="from_data", raw_data={(n, sz): c0})
        return wfn

    def save():
        np.savez_compressed(OUT, **store)
        print("  ->", OUT, os.path.getsize(OUT), "bytes", flush=True)

    # ---- sigma at norb = 14 and 16, both operator classes (bench.py's exact inputs) --------
    for norb in (14, 16):
        if only not in (None, f"sigma{norb}"):
            continue
        n, sz = norb, 0
        na, nb, la, lb = synth.sector_dims(n, sz, norb)
        c0 = synth.state(la, lb, seed=synth.seed_for(norb, 50))
        wfn = wavefunction(n, sz, norb, c0)
        for kind in ("real8", "herm"):
            h1, h2 = synth.integrals(norb, kind)
            t0 = time.perf_counter()
            sigma = wfn.sector((n, sz)).apply((h1, h2)).coeff
            dt = time.perf_counter() - t0
            tag = f"sigma{norb}_{kind}"
            put(store, tag, signature(sigma, 20262000 + norb), meta=[n, sz, norb],
                ref_seconds=[dt], ref_threads=[os.cpu_count()])
            print(tag, "norm", np.linalg.norm(sigma), f"{dt:.1f} s", flush=True)
            del sigma
            save()
        del wfn, c0

    # ---- Taylor time_evolve at norb = 14 (BASELINE config 1) ---------------------------------
    if only in (None, "taylor14"):
        norb, n, sz = 14, 14, 0
        na, nb, la, lb = synth.sector_dims(n, sz, norb)
        h1, h2 = synth.integrals(norb, "real8", scale=0.05)
        e0 = -1.0
        ham = fqe.get_restricted_hamiltonian((h1, h2), e_0=e0)
        c0 = synth.state(la, lb, seed=14)
        wfn = wavefunction(n, sz, norb, c0)
        # |H| from three power iterations, then t = 0.5/|H| (2 significant digits, stored) so
        # the 30-term Taylor cap is not hit
        x = wfn
        for _ in range(3):
            y = x.apply(ham)
            hn = y.norm()
            y.scale(1.0 / hn)
            x = y
        t = float(f"{0.5 / hn:.2g}")
        print("taylor14: |H| ~", hn, "t =", t, flush=True)
        del x, y
        t0 = time.perf_counter()
        ev = wfn.time_evolve(t, ham)
        dt = time.perf_counter() - t0
        cev = ev.get_coeff((n, sz))
        put(store, "taylor14", signature(cev, 20262114), meta=[n, sz, norb], t=[t], e0=[e0],
            scale=[0.05], state_seed=[14], ref_seconds=[dt])
        print("taylor14 norm", np.linalg.norm(cev), f"{dt:.1f} s", flush=True)
        agu = wfn.apply_generated_unitary(t, "taylor", ham).get_coeff((n, sz))
        put(store, "taylor14_agu", signature(agu, 20262115))
        del wfn, ev, cev, agu
        save()

    # ---- DiagonalCoulomb apply + evolve at norb = 16, NON-symmetric v -----------
    if only in (None, "dc16"):
        norb, n, sz = 16, 16, 0
        na, nb, la, lb = synth.sector_dims(n, sz, norb)
        c0 = synth.state(la, lb, seed=synth.seed_for(norb, 50))
        wfn = wavefunction(n, sz, norb, c0)
        vij = synth.diagonal_coulomb_matrix(norb, 3, symmetric=False)
        e0, t = 0.35, 0.1
        dch = fqe.get_diagonalcoulomb_hamiltonian(vij, e_0=e0)
        t0 = time.perf_counter()
        out = wfn.apply(dch).get_coeff((n, sz))
        put(store, "dc16_apply", signature(out, 20262216), meta=[n, sz, norb], e0=[e0], t=[t],
            vij_seed=[3])
        print("dc16 apply", np.linalg.norm(out), f"{time.perf_counter() - t0:.1f} s", flush=True)
        del out
        out = wfn.time_evolve(t, dch).get_coeff((n, sz))
        put(store, "dc16_evolve", signature(out, 20262217))
        print("dc16 evolve", np.linalg.norm(out), flush=True)
        del out
        # 4-index input (diag[k] = h[kkkk], vij = -h[ijij]: diagonal_coulomb.py:54-72)
        rng = np.random.default_rng(20262218)
        h4 = np.zeros((norb,) * 4)
        for i in range(norb):
            for j in range(norb):
                h4[i, j, i, j] = -vij[i, j]
            h4[i, i, i, i] = rng.standard_normal()
        store["dc16_h4diag"] = np.array([h4[i, i, i, i] for i in range(norb)])
        out = wfn.apply(fqe.get_diagonalcoulomb_hamiltonian(h4)).get_coeff((n, sz))
        put(store, "dc16_apply4", signature(out, 20262219))
        del out, wfn
        save()

    # ---- dense 3-body apply, profiling/profile_3_body.py's tensor (BASELINE config 4) --------
    if only in (None, "body3"):
        for norb in (10, 12) if "--body3-12" in sys.argv else (10,):
            n, sz = norb, 0
            na, nb, la, lb = synth.sector_dims(n, sz, norb)
            idx = np.indices((norb,) * 6)
            h3 = ((idx[0] + idx[3]) * (idx[1] + idx[4]) * (idx[2] + idx[5]) * 0.002).astype(
                np.complex128)
            del idx
            h1 = np.zeros((norb,) * 2, dtype=np.complex128)
            h2 = np.zeros((norb,) * 4, dtype=np.complex128)
            c0 = synth.state(la, lb, seed=synth.seed_for(norb, 53))
            wfn = wavefunction(n, sz, norb, c0)
            t0 = time.perf_counter()
            out = wfn.apply((h1, h2, h3)).get_coeff((n, sz))
            dt = time.perf_counter() - t0
            put(store, f"body3_{norb}", signature(out, 20262300 + norb), meta=[n, sz, norb],
                ref_seconds=[dt])
            print(f"body3 norb={norb} norm", np.linalg.norm(out), f"{dt:.1f} s", flush=True)
            del out, wfn
            save()

    # ---- Chebyshev propagator and an orbital rotation (transform) at norb = 12 ---------------
    if only in (None, "cheb12"):
        from scipy.linalg import expm
        norb, n, sz = 12, 12, 0
        na, nb, la, lb = synth.sector_dims(n, sz, norb)
        h1, h2 = synth.integrals(norb, "real8", scale=0.05)
        e0 = 0.25
        ham = fqe.get_restricted_hamiltonian((h1, h2), e_0=e0)
        c0 = synth.state(la, lb, seed=synth.seed_for(norb, 56))
        wfn = wavefunction(n, sz, norb, c0)
        x = wfn
        for _ in range(4):
            y = x.apply(ham)
            hn = y.norm()
            y.scale(1.0 / hn)
            x = y
        lim = [float(f"{-1.2 * hn:.3g}"), float(f"{1.2 * hn:.3g}")]   # generous spectral bounds
        t = float(f"{2.0 / hn:.2g}")
        t0 = time.perf_counter()
        ch = wfn.apply_generated_unitary(t, "chebyshev", ham, spec_lim=lim).get_coeff((n, sz))
        dt = time.perf_counter() - t0
        put(store, "cheb12", signature(ch, 20262512), meta=[n, sz, norb], t=[t], e0=[e0],
            scale=[0.05], spec_lim=lim, ref_seconds=[dt])
        print("cheb12 norm", np.linalg.norm(ch), "lim", lim, "t", t, f"{dt:.1f} s", flush=True)
        rng = np.random.default_rng(20262513)
        a = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
        rot = expm(-0.3j * (a + a.conj().T))
        w2 = wavefunction(n, sz, norb, c0)
        w2.transform(rot)
        put(store, "transform12", signature(w2.get_coeff((n, sz)), 20262514), rot=rot)
        print("transform12 norm", w2.norm(), flush=True)
        save()

    # ---- individual n-body operators at norb = 12 (BASELINE config 5, the sweep) -------------
    if only in (None, "nbody12"):
        norb, n, sz = 12, 12, 0
        na, nb, la, lb = synth.sector_dims(n, sz, norb)
        c0 = synth.state(la, lb, seed=synth.seed_for(norb, 57))
        sec = wavefunction(n, sz, norb, c0).sector((n, sz))
        coeff, tm = 0.3 - 0.2j, 0.41
        ops = [([11, 4, 1], [9, 3, 0], [], []), ([7, 2], [5, 0], [10], [3]),
               ([6], [1], [8, 3], [11, 2]), ([], [], [9, 5, 0], [10, 4, 2]),
               ([5, 3], [5, 3], [7], [7])]
        store["nbody12_ops"] = np.array([repr(o) for o in ops])
        store["nbody12_coeff"], store["nbody12_time"] = np.array([coeff]), np.array([tm])
        store["nbody12_meta"] = np.array([n, sz, norb])
        for k, (da, ua, db, ub) in enumerate(ops):
            out = sec.apply_individual_nbody(coeff, da, ua, db, ub).coeff
            put(store, f"nbody12_apply{k}", signature(out, 20262600 + k))
            if da == ua and db == ub:
                tmp = wavefunction(n, sz, norb, c0.copy()).sector((n, sz))
                tmp.evolve_inplace_individual_nbody_trivial(tm, coeff, da, db)
                ev = tmp.coeff
            else:
                ev = sec.evolve_individual_nbody_nontrivial(tm, coeff, da, ua, db, ub).coeff
            put(store, f"nbody12_evolve{k}", signature(ev, 20262650 + k))
            print(f"nbody12 op {k}: |apply| {np.linalg.norm(out):.6f} |evolve| {np.linalg.norm(ev):.6f}",
                  flush=True)
        save()

    # ---- 1- and 2-particle RDMs (plain and transition) at norb = 10 and 12 ------------------
    if only in (None, "rdm"):
        for norb in (10, 12):
            n, sz = norb, 0
            na, nb, la, lb = synth.sector_dims(n, sz, norb)
            ket = synth.state(la, lb, seed=synth.seed_for(norb, 54))
            bra = synth.state(la, lb, seed=synth.seed_for(norb, 55))
            ksec = wavefunction(n, sz, norb, ket).sector((n, sz))
            bsec = wavefunction(n, sz, norb, bra).sector((n, sz))
            t0 = time.perf_counter()
            r1, r2 = ksec.rdm12()
            t1, t2 = ksec.rdm12(bsec)
            dt = time.perf_counter() - t0
            rng = np.random.default_rng(20262400 + norb)
            pick = rng.choice(norb ** 4, size=4096, replace=False)
            for tag, a1, a2 in ((f"rdm{norb}", r1, r2), (f"trdm{norb}", t1, t2)):
                store[f"{tag}_1"] = np.asarray(a1)
                store[f"{tag}_2_idx"] = pick
                store[f"{tag}_2_val"] = np.asarray(a2).reshape(-1)[pick]
                store[f"{tag}_2_norm"] = np.array([np.linalg.norm(a2)])
                store[f"{tag}_2_trace"] = np.array([np.einsum("ijij", a2)])
            store[f"rdm{norb}_meta"] = np.array([n, sz, norb])
            print(f"rdm12 norb={norb}: trace(rdm1) {np.trace(r1):.12f}  {dt:.1f} s", flush=True)
            save()


if __name__ == "__main__":
    main()
