"""Sliced (INT8 tcgen05) contraction against the FP64 DMMA paths: relative difference of sigma,
which path ran, and timings.  python scripts/ozaki_check.py [max_norb]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import numpy as np
import torch
import fqe_b200 as fqe
from fqe_b200 import synth, lib as L
from fqe_b200.fqe_data import FqeData, DenseOperator, release_workspace

lib = L.load()
PATH = {0: "none", 1: "three-kernel", 2: "fused-dmma", 3: "sliced-i8"}
max_norb = int(sys.argv[1]) if len(sys.argv) > 1 else 12


def run(na, nb, norb, kind, env):
    for k in ("FQEB_OZAKI", "FQEB_FUSION"):
        os.environ.pop(k, None)
    os.environ.update(env)
    h1, h2 = synth.integrals(norb, kind)
    d = FqeData(na, nb, norb)
    d.set_wfn(strategy="from_data", raw_data=synth.state(d.lena(), d.lenb(), seed=synth.seed_for(norb, 50)))
    op = DenseOperator(norb, h1, h2)
    out = d.apply_operator(op)
    torch.cuda.synchronize()
    path = lib.fqeb_sigma_last_path()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        out = d.apply_operator(op)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return out, path, dt


cases = [(1, 1, 2), (2, 1, 4), (2, 2, 4), (2, 3, 6), (3, 3, 6), (4, 3, 7), (4, 4, 8), (0, 2, 4),
         (3, 0, 5), (4, 4, 4), (1, 6, 8), (5, 5, 10), (6, 6, 12), (7, 7, 14), (8, 8, 16)]
for na, nb, norb in cases:
    if norb > max_norb:
        continue
    for kind in ("real8", "general_real"):
        k = "real8" if kind == "real8" else "herm"
        if kind == "general_real":
            # non-symmetric real operator: full pair space norb^2 (sliced path if norb^2 <= 144)
            rng = np.random.default_rng(5 + norb)
            h1 = rng.standard_normal((norb, norb))
            h2 = 0.1 * rng.standard_normal((norb,) * 4)
            synth_int = synth.integrals
            synth.integrals = lambda n, kk, seed=None, scale=0.1: (h1.astype(complex), h2.astype(complex))
        try:
            ref, pref, tref = run(na, nb, norb, k, {"FQEB_OZAKI": "0"})
            out, pout, tout = run(na, nb, norb, k, {})
        finally:
            if kind == "general_real":
                synth.integrals = synth_int
        err = float((torch.linalg.norm(out - ref) / torch.linalg.norm(ref)).item()) if ref.numel() else 0.0
        print(f"({na},{nb},{norb}) {kind:12s} ref={PATH[pref]:12s} {tref*1e3:9.3f} ms | "
              f"default={PATH[pout]:12s} {tout*1e3:9.3f} ms | rel diff {err:.2e}", flush=True)
    release_workspace()
