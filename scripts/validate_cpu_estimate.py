"""One FULL reference sigma build at a given norb on the host cores, next to the
exact-work-sample estimate that bench.py's cpu_baseline uses (oracle/ref_harness.py)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
from oracle import ref_harness as R
from fqe_b200 import synth
norb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
na = nb = norb // 2
g = R.graph(na, nb, norb)
h1, h2 = synth.integrals(norb, "real8")
c = synth.state(g.lena, g.lenb, seed=synth.seed_for(norb, 50))
est, desc = R.estimate_sigma_seconds(g, c, h1, h2, 15.0)
t0 = time.perf_counter()
R.sigma_restricted(g, c, h1, h2)
full = time.perf_counter() - t0
print(json.dumps({"norb": norb, "cores": len(os.sched_getaffinity(0)), "full_seconds": full,
                  "estimate_seconds": est, "ratio": est / full, "sample": desc}))
