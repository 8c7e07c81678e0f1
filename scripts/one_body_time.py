"""Where the time of a one-body sigma at norb=16 goes; Taylor terms at norb=14 and the path they take."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import numpy as np, torch
import fqe_b200 as fqe
from fqe_b200 import synth, lib as L
from fqe_b200.fqe_data import DenseOperator, FqeData
lib = L.load()
def ev_time(fn, reps=3, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
norb = 16
d = FqeData(8, 8, norb)
d.set_wfn(strategy="from_data", raw_data=synth.state(d.lena(), d.lenb(), seed=1))
rng = np.random.default_rng(3); k = rng.standard_normal((norb, norb)); k = 0.5 * (k + k.T)
op = DenseOperator(norb, k, None)
out = torch.empty_like(d.coeff)
print("one-body sigma norb=16, prepared operator, out given: %.2f ms" % ev_time(lambda: d.apply_operator(op, out=out)))
print("one-body sigma norb=16, prepared operator:            %.2f ms" % ev_time(lambda: d.apply_operator(op)))
print("one-body sigma norb=16, FqeData.apply((h1,)):          %.2f ms" % ev_time(lambda: d.apply((k,))))
t0 = time.perf_counter(); DenseOperator(norb, k, None); torch.cuda.synchronize(); print("operator creation: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
# Taylor at norb=14
norb = 14
w = fqe.Wavefunction([[14, 0, 14]])
w.set_wfn(strategy="from_data", raw_data={(14, 0): synth.state(3432, 3432, seed=2)})
h1, h2 = synth.integrals(norb, "real8")
ham = fqe.get_restricted_hamiltonian((h1, h2))
sec = w.sector((14, 0))
op2 = DenseOperator(norb, h1, h2)
print("sigma norb=14: %.2f ms, path %d" % (ev_time(lambda: sec.apply_operator(op2)), lib.fqeb_sigma_last_path()))
opt = DenseOperator(norb, -0.0084j * h1, -0.0084j * h2)
print("sigma norb=14 with -i t H: %.2f ms, path %d, kind %d" % (ev_time(lambda: sec.apply_operator(opt)), lib.fqeb_sigma_last_path(), opt.kind))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = w.time_evolve(0.0084, ham)
    torch.cuda.synchronize()
    print("time_evolve norb=14: %.1f ms, %d terms, last path %d" % ((time.perf_counter() - t0) * 1e3, getattr(w, "last_expansion_order", -1), lib.fqeb_sigma_last_path()))
import copy
s2 = copy.deepcopy(sec)
torch.cuda.synchronize(); t0 = time.perf_counter()
n = s2.taylor_inplace(opt)
torch.cuda.synchronize()
print("taylor_inplace norb=14: %.1f ms, %d terms" % ((time.perf_counter() - t0) * 1e3, n))
