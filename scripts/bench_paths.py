"""Secondary measurements for the other BASELINE.json configs (not the bench.py contract):
  config 2: norb=14 sigma apply + Taylor time_evolve on one GPU
  config 3: norb=16 DiagonalCoulomb apply / evolve (HBM-bound: 32*L^2 bytes per pass)
  BLAS-1 : axpy / scale / norm / fused axpy+norm at norb=16
Prints one JSON object; CUDA-event timing, 3 warm-ups, inputs larger than L2."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import numpy as np
import torch
import fqe_b200 as fqe
from fqe_b200 import synth

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"hbm_peak_GBps": HBM}

# ---- config 3: diagonal Coulomb at norb=16 ------------------------------------------------
norb, n, sz = 16, 16, 0
na, nb, la, lb = synth.sector_dims(n, sz, norb)
wfn = fqe.Wavefunction([[n, sz, norb]])
gen = torch.Generator(device="cuda").manual_seed(7)
wfn.set_wfn(strategy="from_data", raw_data={(n, sz): torch.view_as_complex(
    torch.randn((la, lb, 2), dtype=torch.float64, device="cuda", generator=gen))})
wfn.normalize()
sec = wfn.sector((n, sz))
vij = synth.diagonal_coulomb_matrix(norb, 3)
diag = np.zeros(norb)
nbytes = 32.0 * la * lb
ms = timed(lambda: sec.apply_diagonal_coulomb(diag, vij, inplace=True))
out["dc_apply_norb16"] = {"ms": ms, "GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / HBM}
ms = timed(lambda: sec.evolve_diagonal_coulomb(-0.1j * diag, -0.1j * vij, inplace=True))
out["dc_evolve_norb16"] = {"ms": ms, "GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / HBM}
other = sec.empty_copy(zero=False)
other.coeff.copy_(sec.coeff)
ms = timed(lambda: sec.ax_plus_y(1e-6 + 1e-6j, other))
out["zaxpy_norb16"] = {"ms": ms, "GBps": 48.0 * la * lb / ms / 1e6}
ms = timed(lambda: sec.scale(1.0 + 1e-9j))
out["zscal_norb16"] = {"ms": ms, "GBps": nbytes / ms / 1e6}
ms = timed(lambda: sec.norm())
out["znorm_norb16"] = {"ms": ms, "GBps": 16.0 * la * lb / ms / 1e6}
ms = timed(lambda: sec.axpy_norm(1e-6, other))
out["axpy_norm_norb16"] = {"ms": ms, "GBps": 48.0 * la * lb / ms / 1e6}

# ---- config 3: one Trotter step of a double-factorised H = orbital rotation o DC evolve ------
rng = np.random.default_rng(3)
kmat = rng.standard_normal((norb, norb))
kmat = 0.5 * (kmat + kmat.T)
rot_ham = fqe.get_restricted_hamiltonian((kmat,))
dc_ham = fqe.get_diagonalcoulomb_hamiltonian(vij)
ms = timed(lambda: sec.apply((kmat,)), reps=3, warm=1)
out["one_body_sigma_norb16"] = {"ms": ms}
wfn.time_evolve(0.01, rot_ham)          # warm-up: builds the occupancy lists once
torch.cuda.synchronize()
t0 = time.perf_counter()
step = wfn.time_evolve(0.01, rot_ham)
step = step.time_evolve(0.01, dc_ham, inplace=True)
torch.cuda.synchronize()
norm0 = wfn.norm()
out["trotter_step_norb16"] = {"seconds": time.perf_counter() - t0,
                              "route": "transform -> evolve_diagonal -> transform, then DC evolve",
                              "norm_ratio_after": step.norm() / norm0}
# one orbital rotation alone (Wavefunction.transform: 2*norb column passes), in place
from scipy.linalg import expm
umat = expm(-0.3j * kmat)
rot = wfn.time_evolve(0.0, dc_ham)      # a copy to rotate
torch.cuda.synchronize()
t0 = time.perf_counter()
rot.transform(umat)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
# algorithmic bytes per rotation: 2*norb columns x 16*L^2*(n/2 + 2)
rbytes = 2 * norb * 16.0 * la * lb * (na / 2 + 2)
out["transform_norb16"] = {"seconds": dt, "algorithmic_GBps": rbytes / dt / 1e9,
                           "frac_hbm": rbytes / dt / 1e9 / HBM,
                           "norm_ratio_after": rot.norm() / norm0}
del rot
del wfn, sec, other, step
torch.cuda.empty_cache()

# ---- config 2: norb=14 sigma + Taylor time_evolve -----------------------------------------
norb, n, sz = 14, 14, 0
na, nb, la, lb = synth.sector_dims(n, sz, norb)
h1, h2 = synth.integrals(norb, "real8", scale=0.05)
ham = fqe.get_restricted_hamiltonian((h1, h2), e_0=-1.0)
wfn = fqe.Wavefunction([[n, sz, norb]])
wfn.set_wfn(strategy="from_data", raw_data={(n, sz): synth.state(la, lb, seed=14)})
ms = timed(lambda: wfn.apply(ham), reps=3, warm=2)
out["sigma_norb14_real8"] = {"ms": ms, "sigma_per_s": 1e3 / ms}
# |H| from a few power iterations, then t = 0.5/|H| so Taylor converges well under 30 terms
x = wfn
for _ in range(3):
    y = x.apply(ham)
    hn = y.norm()
    y.scale(1.0 / hn)
    x = y
t = 0.5 / hn
torch.cuda.synchronize()
t0 = time.perf_counter()
ev = wfn.time_evolve(t, ham)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["time_evolve_norb14"] = {"t": t, "H_norm_est": hn, "taylor_terms": wfn.last_expansion_order
                             if hasattr(wfn, "last_expansion_order") else None,
                             "seconds": dt, "norm_after": ev.norm()}
del wfn, ev, x, y
torch.cuda.empty_cache()

# ---- config 5: sweep of individual 3-body operators at norb=12, n=12 -----------------------
norb, n, sz = 12, 12, 0
na, nb, la, lb = synth.sector_dims(n, sz, norb)
wfn = fqe.Wavefunction([[n, sz, norb]])
wfn.set_wfn(strategy="from_data", raw_data={(n, sz): synth.state(la, lb, seed=12)})
sec = wfn.sector((n, sz))
rng = np.random.default_rng(5)
terms = []
for _ in range(200):
    ka = int(rng.integers(0, 4))                     # alpha operators of the 3-body term
    daga = sorted(rng.choice(norb, ka, replace=False).tolist(), reverse=True)
    unda = sorted(rng.choice(norb, ka, replace=False).tolist(), reverse=True)
    dagb = sorted(rng.choice(norb, 3 - ka, replace=False).tolist(), reverse=True)
    undb = sorted(rng.choice(norb, 3 - ka, replace=False).tolist(), reverse=True)
    terms.append((complex(rng.standard_normal(), rng.standard_normal()), daga, unda, dagb, undb))
acc = sec.empty_copy(zero=True)


def sweep():
    for c, da, ua, db, ub in terms:
        acc.apply_individual_nbody_accumulate(c, sec, da, ua, db, ub)


ms = timed(sweep, reps=3, warm=1)
out["individual_3body_sweep_norb12"] = {"terms": len(terms), "ms_per_term": ms / len(terms)}
ev_terms = [t for t in terms if not (t[1] == t[2] and t[3] == t[4])][:50]


def evolve_sweep():
    cur = sec
    for c, da, ua, db, ub in ev_terms:
        cur = cur.evolve_individual_nbody_nontrivial(0.1, c, da, ua, db, ub)
    return cur


ms = timed(evolve_sweep, reps=2, warm=1)
out["individual_3body_evolve_norb12"] = {"terms": len(ev_terms), "ms_per_term": ms / len(ev_terms),
                                         "norm_after": evolve_sweep().norm()}
print(json.dumps(out))
