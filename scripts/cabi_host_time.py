"""Time fqeb_sigma_restricted_host (pageable numpy arrays in and out) at norb (default 16):
the host-buffer C-ABI call of INTEGRATION.md.  FQEB_COPY_THREADS sets the copy-thread count."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import numpy as np
from fqe_b200 import synth, lib as L
from fqe_b200.fqe_data import fold_restricted
from math import comb

lib = L.load()
norb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
na = nb = norb // 2
la = comb(norb, na)
h1, h2 = synth.integrals(norb, "real8")
h1p, h2p = fold_restricted(h1, h2)
c = np.array(synth.state(la, la, seed=synth.seed_for(norb, 50)), copy=True)
s = np.empty_like(c)

def step():
    rc = lib.fqeb_sigma_restricted_host(norb, na, nb, h1p.ctypes.data, h2p.ctypes.data,
                                        c.ctypes.data, s.ctypes.data)
    if rc != 0:
        raise RuntimeError(lib.fqeb_last_error().decode())

step(); step()
t0 = time.perf_counter()
for _ in range(reps):
    step()
dt = (time.perf_counter() - t0) / reps
print(f"norb={norb} copy threads={os.environ.get('FQEB_COPY_THREADS', 'default')}: "
      f"{dt * 1e3:.1f} ms per host-buffer sigma ({1 / dt:.3f} sigma/s), |sigma|={np.linalg.norm(s):.6f}")
