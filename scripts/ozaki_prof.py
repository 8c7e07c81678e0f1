"""Phase times and in-kernel cycle counters of the sliced contraction at norb (default 16).
FQEB_OZAKI_PROF=1 python scripts/ozaki_prof.py [norb]"""
import ctypes, os, sys, time
os.environ.setdefault("FQEB_OZAKI_PROF", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import numpy as np
import torch
from fqe_b200 import synth, lib as L
from fqe_b200.fqe_data import FqeData, DenseOperator

lib = L.load()
norb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
na = nb = norb // 2
h1, h2 = synth.integrals(norb, "real8")
d = FqeData(na, nb, norb)
d.set_wfn(strategy="from_data", raw_data=synth.state(d.lena(), d.lenb(), seed=synth.seed_for(norb, 50)))
op = DenseOperator(norb, h1, h2)
out = d.apply_operator(op)
torch.cuda.synchronize()
prof = (ctypes.c_uint64 * 8)()
lib.fqeb_ozaki_profile(prof)      # reset
lib.fqeb_profile_enable(1)
ms3, cnt3 = (ctypes.c_double * 3)(), (ctypes.c_int64 * 3)()
lib.fqeb_profile_collect(ms3, cnt3)
reps = 3
t0 = time.perf_counter()
for _ in range(reps):
    out = d.apply_operator(op)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
lib.fqeb_profile_collect(ms3, cnt3)
lib.fqeb_ozaki_profile(prof)
ntiles = ((d.lena() + 7) // 8) * ((d.lenb() + 7) // 8) * reps   # 8 x 8 determinant tiles (chunk edges ignored)
print(f"norb={norb} path={lib.fqeb_sigma_last_path()} sigma {dt*1e3:.1f} ms; phases per sigma (ms): "
      f"prepass {ms3[0]/reps:.2f} contract {ms3[1]/reps:.2f} scatter {ms3[2]/reps:.2f}; launches {list(cnt3)}")
names = ["issuer total", "issuer wait tile", "issuer wait slot", "producer wait staging", "producer produce",
         "drain (incl. waiting for MMAs)", "drain store"]
nctas = prof[7]
if nctas:
    tiles_per_cta = ntiles / nctas
    for i, nm in enumerate(names):
        print(f"  {nm:32s} {prof[i]/nctas/tiles_per_cta:10.0f} cycles per tile per CTA")
    print(f"  ({nctas} CTA launches, {tiles_per_cta:.1f} tiles per CTA)")
