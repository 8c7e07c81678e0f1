"""Per-phase times of one norb=16 sigma build (library-internal events), for A/B runs of
kernel options set through environment variables."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import torch
import fqe_b200 as fqe
from fqe_b200 import lib as L, synth
from fqe_b200.fqe_data import DenseOperator

norb = int(sys.argv[sys.argv.index("--norb") + 1]) if "--norb" in sys.argv else 16
na, nb, la, lb = synth.sector_dims(norb, 0, norb)
h1, h2 = synth.integrals(norb, "real8")
wfn = fqe.Wavefunction([[norb, 0, norb]])
sec = wfn.sector((norb, 0))
sec.set_wfn(strategy="from_data", raw_data=synth.state(la, lb, seed=1))
op = DenseOperator(norb, h1, h2)
lib = L.load()
for _ in range(2):
    sec.apply_operator(op)
torch.cuda.synchronize()
lib.fqeb_profile_enable(1)
ms3, cnt3 = (ctypes.c_double * 3)(), (ctypes.c_int64 * 3)()
lib.fqeb_profile_collect(ms3, cnt3)
reps = 3
for _ in range(reps):
    s = sec.apply_operator(op)
torch.cuda.synchronize()
lib.fqeb_profile_collect(ms3, cnt3)
print({k: os.environ[k] for k in os.environ if k.startswith("FQEB_")},
      "gather %.1f contract %.1f scatter %.1f ms" % tuple(x / reps for x in ms3),
      "checksum %.6f" % float(torch.view_as_real(s).abs().sum().item()))
