"""One device-resident sigma build at norb=16 (or --norb N): the smallest program that
launches the three sigma kernels, for ncu captures.

    ncu --set full --clock-control none -k regex:k_make_coeff -c 1 -o out python scripts/one_sigma.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import torch
import fqe_b200 as fqe
from fqe_b200 import synth
from fqe_b200.fqe_data import DenseOperator

norb = int(sys.argv[sys.argv.index("--norb") + 1]) if "--norb" in sys.argv else 16
kind = sys.argv[sys.argv.index("--kind") + 1] if "--kind" in sys.argv else "real8"
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 1
na, nb, la, lb = synth.sector_dims(norb, 0, norb)
h1, h2 = synth.integrals(norb, kind)
wfn = fqe.Wavefunction([[norb, 0, norb]])
sec = wfn.sector((norb, 0))
sec.set_wfn(strategy="from_data", raw_data=synth.state(la, lb, seed=1))
op = DenseOperator(norb, h1, h2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sigma = sec.apply_operator(op)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    sigma = sec.apply_operator(op)
e1.record()
torch.cuda.synchronize()
print("norb", norb, kind, "ms per sigma", e0.elapsed_time(e1) / reps,
      "checksum", float(torch.view_as_real(sigma).abs().sum().item()))
