"""Throughput of fqeb_gram_accumulate (the RDM reduction kernel) on rdm12-like shapes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
import torch
from fqe_b200 import lib as L

lib = L.load()
st = torch.cuda.current_stream().cuda_stream
for (m, ncols) in [(144, 853776), (144, 853776), (196, 3432 * 1200), (256, 12870 * 81)]:
    n = m + 1
    bra = torch.randn(m, ncols, 2, device="cuda", dtype=torch.float64)
    last = torch.randn(ncols, 2, device="cuda", dtype=torch.float64)
    g = torch.zeros(m, n, 2, device="cuda", dtype=torch.float64)
    def run():
        L.call("fqeb_gram_accumulate", m, n, ncols, bra.data_ptr(), ncols, bra.data_ptr(), ncols,
               last.data_ptr(), g.data_ptr(), st)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flop = 8.0 * m * n * ncols
    braz = torch.view_as_complex(bra)
    t0 = time.perf_counter(); ref = braz.conj() @ braz.transpose(0, 1); torch.cuda.synchronize()
    e0.record(); ref = braz.conj() @ braz.transpose(0, 1); e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1)
    ket2 = bra.clone()
    def run2():
        L.call("fqeb_gram_accumulate", m, n, ncols, bra.data_ptr(), ncols, ket2.data_ptr(), ncols,
               last.data_ptr(), g.data_ptr(), st)
    run2(); run2(); torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        run2()
    e1.record(); torch.cuda.synchronize()
    print(f"   transition form (bra != ket): {e0.elapsed_time(e1) / 3:.2f} ms = {flop / (e0.elapsed_time(e1) / 3) / 1e9:.1f} TFLOP/s")
    print(f"M={m} N={n} ncols={ncols}: k_gram {ms:.2f} ms = {flop / ms / 1e9:.1f} TFLOP/s; "
          f"torch (cuBLAS zgemm, N={m}) {ms_ref:.2f} ms = {8.0 * m * m * ncols / ms_ref / 1e9:.1f} TFLOP/s")
