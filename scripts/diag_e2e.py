"""Diagnostic: wall-clock breakdown of one end-to-end sigma step (H2D, operator
preparation, sigma build, D2H) at a given norb.  Not part of the bench contract."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import torch
import fqe_b200 as fqe
from fqe_b200 import synth

norb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, sz = norb, 0
na, nb, la, lb = synth.sector_dims(n, sz, norb)
h1, h2 = synth.integrals(norb, "real8")
host_c = torch.from_numpy(synth.state(la, lb, seed=1)).pin_memory()
host_out = torch.empty((la, lb), dtype=torch.complex128).pin_memory()
wfn = fqe.Wavefunction([[n, sz, norb]])
sector = wfn.sector((n, sz))

def T():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(3):
    t0 = T(); sector.coeff.copy_(host_c, non_blocking=True)
    t1 = T(); ham = fqe.get_restricted_hamiltonian((h1, h2)); op = wfn._dense_operator(ham.tensors())
    t2 = T(); out = sector.apply_operator(op)
    t3 = T(); host_out.copy_(out, non_blocking=True)
    t4 = T(); del op
    t5 = T()
    print(f"rep {rep}: h2d {t1-t0:.3f}  op_create {t2-t1:.3f}  sigma {t3-t2:.3f}  d2h {t4-t3:.3f}  op_destroy {t5-t4:.3f}")
