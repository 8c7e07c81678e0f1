"""Stress test for the contraction kernel: repeat many small launches and compare with
torch.matmul; prints the number of mismatches per configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import numpy as np, torch
from fqe_b200 import lib as L
from fqe_b200.fqe_data import DenseOperator

lib = L.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(0)
for norb, kind in [(6, "real"), (6, "complex"), (4, "real"), (8, "real"), (12, "real"), (16, "real"),
                   (16, "complex")]:
    npair = norb * norb
    h2p = rng.standard_normal((npair, npair)) + 1j * rng.standard_normal((npair, npair))
    if kind == "real":
        h2p = h2p.real.astype(np.complex128)
    op = DenseOperator(norb, np.zeros((norb, norb)), -np.moveaxis(h2p.reshape((norb,) * 4), 2, 1))
    for ncols in (1, 300, 5000):
        ld = ((ncols + 127) // 128) * 128
        drows = lib.fqeb_contract_dvec_rows(op.handle, npair)
        bad = 0
        worst = 0.0
        for r in range(reps):
            dv = torch.zeros((drows, ld), dtype=torch.complex128, device="cuda")
            dh = torch.randn((npair, ncols), dtype=torch.complex128, device="cuda")
            dv[:npair, :ncols] = dh
            ev = torch.zeros((npair + 8, ld), dtype=torch.complex128, device="cuda")
            L.call("fqeb_contract", op.handle, dv.data_ptr(), ld, ev.data_ptr(), ld, ncols, 0, npair,
                   None)
            ref = torch.from_numpy(h2p).cuda() @ dh
            err = float(torch.linalg.norm(ev[:npair, :ncols] - ref) / torch.linalg.norm(ref))
            worst = max(worst, err)
            bad += err > 1e-12
        print(f"norb={norb:2d} {kind:7s} ncols={ncols:5d}: {bad}/{reps} mismatches, worst rel err {worst:.2e}")
