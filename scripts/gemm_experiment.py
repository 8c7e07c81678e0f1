"""Time the contraction kernel alone on the norb=16 shapes (A/B harness for kernel variants).

    FQEB_B200_LIB=/path/to/variant.so python scripts/gemm_experiment.py [rows]

Prints TFLOP/s for the real pair-symmetric (P=136) and the complex (P=256) operator on a
D chunk of `rows` alpha rows (default 1839, the chunk bench.py runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import numpy as np, torch
from fqe_b200 import lib as L, synth
from fqe_b200.fqe_data import DenseOperator

lib = L.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1839
norb, lenb = 16, 12870
out = {}
for kind in ("real8", "herm"):
    h1, h2 = synth.integrals(norb, kind)
    op = DenseOperator(norb, h1, h2)
    P = op.npair
    r = rows if kind == "real8" else rows // 2
    ncols = r * lenb
    ld = ((ncols + 127) // 128) * 128
    drows = lib.fqeb_contract_dvec_rows(op.handle, P)
    dv = torch.zeros((drows, ld), dtype=torch.complex128, device="cuda")
    dv[:P].normal_()
    ev = torch.empty((P + 8, ld), dtype=torch.complex128, device="cuda")
    def run():
        L.call("fqeb_contract", op.handle, dv.data_ptr(), ld, ev.data_ptr(), ld, ncols, 0, P, None)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = (4.0 if kind == "real8" else 8.0) * P * P * ncols
    out[kind] = (ms, flops / ms / 1e9)
    del dv, ev
    torch.cuda.empty_cache()
print(os.environ.get("FQEB_B200_LIB", "product"), {k: "%.1f ms %.2f TF" % v for k, v in out.items()})
