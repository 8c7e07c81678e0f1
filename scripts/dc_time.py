"""Time the diagonal-Coulomb apply / evolve kernels at norb=16 (A/B harness)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import numpy as np, torch
import fqe_b200 as fqe
from fqe_b200 import synth
norb, n, sz = 16, 16, 0
na, nb, la, lb = synth.sector_dims(n, sz, norb)
wfn = fqe.Wavefunction([[n, sz, norb]])
wfn.set_wfn(strategy="from_data", raw_data={(n, sz): torch.view_as_complex(
    torch.randn((la, lb, 2), dtype=torch.float64, device="cuda"))})
sec = wfn.sector((n, sz))
vij = synth.diagonal_coulomb_matrix(norb, 3)
diag = np.zeros(norb)
def timed(fn, reps=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
a = timed(lambda: sec.apply_diagonal_coulomb(diag, 1e-3 * vij, inplace=True))
e = timed(lambda: sec.evolve_diagonal_coulomb(-0.1j * diag, -0.1j * vij, inplace=True))
nbytes = 32.0 * la * lb
print(os.environ.get("FQEB_B200_LIB", "product"), "apply %.3f ms %.0f GB/s  evolve %.3f ms %.0f GB/s" % (a, nbytes / a / 1e6, e, nbytes / e / 1e6))
