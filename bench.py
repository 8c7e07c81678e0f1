#!/usr/bin/env python
"""Benchmark: sigma applies / s for a random RestrictedHamiltonian at
(norb=16, n=16, Sz=0) -- BASELINE.json's metric -- on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps K --warmup W     (CPU arm)

A "step" is one sigma build  sigma = H C  over the whole determinant space
(L^2 = 165 636 900 determinants at norb=16).  One JSON line is printed by rank 0.

* value      device-timed (CUDA events, max over ranks), C already resident in HBM.
* e2e        same metric through the public API with HOST buffers: every step copies C
             from pinned host memory to the device, prepares the operator, builds
             sigma and copies sigma back to pinned host memory.
* roofline   the dominant kernel is the FP64 tensor-core contraction (k_dgemm):
             achieved = algorithmic flops of the contraction / its CUDA-event time,
             measured live by the library's per-phase events; peak = FP64 GEMM
             throughput of cuBLAS measured in this run (MEASURED_PEAKS.json has no
             FP64 entry).  The HBM-bound gather / scatter phases are
             reported against MEASURED_PEAKS.json's hbm_gbs in "phases".
* cpu_baseline  the UNMODIFIED reference C kernels (oracle/_ref, compiled from
             /root/reference) timed on this host's cores on exact-work samples of the
             same norb=16 workload.
Multi-GPU: C replicated, work sharded (--shard det|pair), one NCCL allreduce of
sigma per step; total work is fixed, so scaling is "strong".
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "openfermion-fqe_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "sigma applies/sec (norb=16,n=16,Sz=0)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--norb", type=int, default=16)
    ap.add_argument("--kind", default="real8", choices=["real8", "herm"],
                    help="real8: real 8-fold symmetric integrals passed as complex128 "
                    "(SURVEY 8d primary recipe); herm: general complex-Hermitian")
    ap.add_argument("--shard", default="det", choices=["det", "pair"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the second operator class (complex-Hermitian) leg")
    ap.add_argument("--verify", action="store_true",
                    help="also build sigma unsharded on every rank and report the relative "
                    "difference to the sharded + allreduced result")
    ap.add_argument("--cpu-budget", type=float, default=15.0,
                    help="seconds of CPU work per reference sample")
    ap.add_argument("--ref-budget", type=float, default=300.0,
                    help="--impl reference: seconds of full sigma builds (at least one is run)")
    return ap.parse_args()


def metric_name(norb):
    return METRIC if norb == 16 else f"sigma applies/sec (norb={norb},n={norb},Sz=0)"


def workload_name(norb, kind):
    cls = ("real 8-fold-symmetric integrals passed as complex128" if kind == "real8" else
           "general complex-Hermitian integrals")
    return (f"RestrictedHamiltonian sigma apply norb={norb} n={norb} Sz=0, random {cls}, "
            f"seed=20260000+100*norb")


# ----------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown",
                                      "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------
# CPU arm: the reference's own C kernels on the host cores
# ----------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads():
    """Give the reference's OpenMP loops every core this process may run on, whatever the
    launcher exported (torchrun sets OMP_NUM_THREADS=1 for its workers), and keep BLAS from
    nesting threads inside them.  Returns the OpenMP thread count actually in effect."""
    import ctypes
    cores = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    # BLAS (scipy's zaxpy, called inside the OpenMP loops) must not nest its own pthreads.  Do NOT
    # set MKL_NUM_THREADS here: with an OpenMP-threaded BLAS that variable caps the whole
    # process's OpenMP team, and the reference's loops then run on ONE thread (measured: cpu
    # time == wall time) while omp_get_max_threads() still reports all cores.
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.pop("MKL_NUM_THREADS", None)
    from oracle import ref_harness as R
    R.lib()                       # loads oracle/_ref/libfqe_ref*.so (and libgomp with it)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(cores)   # libgomp may have read the environment earlier
        gomp.omp_get_max_threads.restype = ctypes.c_int
        return int(gomp.omp_get_max_threads())
    except OSError:
        return None


def cpu_inputs(norb, kind):
    from oracle import ref_harness as R
    from fqe_b200 import synth
    na = nb = norb // 2
    g = R.graph(na, nb, norb)
    h1, h2 = synth.integrals(norb, kind)
    c = synth.state(g.lena, g.lenb, seed=synth.seed_for(norb, 50))
    return R, g, c, h1, h2


def cpu_sigma_seconds(norb, kind, budget_s):
    """Bounded exact-work sample of one reference sigma, extrapolated (ref_harness)."""
    R, g, c, h1, h2 = cpu_inputs(norb, kind)
    return R.estimate_sigma_seconds(g, c, h1, h2, budget_s)


def config_dict(norb, kind, shard, world, op_class=None):
    """The same `config` keys in both arms (the driver compares them)."""
    from math import comb
    la = comb(norb, norb // 2)
    cfg = {
        "workload": workload_name(norb, kind), "norb": norb, "kind": kind,
        "determinants": la * la, "shard": shard, "parallelism": f"{shard}{world}",
        "l2": "inputs larger than L2 (C = %.2f GB, D/E chunks stream from HBM)" %
              (la * la * 16 / 1e9),
        "operator_class": op_class or ("real, pair-symmetric" if kind == "real8" else "complex"),
    }
    return cfg


def run_reference(args):
    """The reference's own C implementation of the path (oracle/_ref, compiled from
    /root/reference/src/fqe/lib/*.c) on all host cores.  Every timed step is ONE FULL sigma
    build at the benchmark size; as many of the requested steps as fit into --ref-budget
    seconds are run (at least one) and `steps` reports how many did."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    threads = use_all_host_threads()
    R, g, c, h1, h2 = cpu_inputs(args.norb, args.kind)
    # warm-up: a small exact-work slice (loads the library, touches the tables and C)
    R.time_sigma_sample(g, c, h1, h2, 16)
    import resource
    times = []
    t_begin = time.perf_counter()
    cpu_begin = resource.getrusage(resource.RUSAGE_SELF)
    sig = None
    while len(times) < max(1, args.steps):
        if times and (time.perf_counter() - t_begin) + times[-1] > args.ref_budget:
            break
        t0 = time.perf_counter()
        sig = R.sigma_restricted(g, c, h1, h2)
        times.append(time.perf_counter() - t0)
    cpu_end = resource.getrusage(resource.RUSAGE_SELF)
    # cores actually kept busy by the timed builds: process CPU time / wall time
    busy = ((cpu_end.ru_utime + cpu_end.ru_stime) - (cpu_begin.ru_utime + cpu_begin.ru_stime)) \
        / max(sum(times), 1e-9)
    mean_s = sum(times) / len(times)
    value = 1.0 / mean_s
    verify = verify_against_golden(sig, args.norb, args.kind, numpy_state=True)
    del sig
    est_s, est_desc = R.estimate_sigma_seconds(g, c, h1, h2, args.cpu_budget)
    sample = ("%d full sigma build(s) of the whole workload, lm_apply_array12_same_spin_opt x2 + "
              "lm_apply_array12_diff_spin_opt as FqeData._apply_array_spatial12_lm calls them; "
              "OpenMP threads in effect = %s on %d usable cores; scipy BLAS zaxpy as in the "
              "reference's Cython shim, BLAS itself single-threaded inside the OpenMP loops"
              % (len(times), threads, cores))
    line = {
        "impl": "reference",
        "metric": metric_name(args.norb), "value": value, "unit": "sigma/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": 0,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * mean_s, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.norb, args.kind, args.shard, args.gpus),
        "cpu_baseline": {"value": value, "unit": "sigma/s", "cores": cores,
                         "omp_threads": threads, "kind": "reference", "sample": sample,
                         "extrapolated": False, "step_seconds": times,
                         "busy_cores": busy},
        "sampled_estimate": {"value": 1.0 / est_s, "unit": "sigma/s", "extrapolated": True,
                             "sample": est_desc},
        "verify_rel_err": verify,
        "e2e": {"value": value, "unit": "sigma/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    if threads is not None and threads != cores:
        line["cpu_baseline"]["warning"] = "OpenMP threads != usable cores"
    if cores > 1 and busy < 1.5:
        line["cpu_baseline"]["warning"] = ("the reference ran on ~%.1f cores although %d are "
                                           "usable: this line under-states it" % (busy, cores))
    emit(line)


def verify_against_golden(state, norb, kind, numpy_state=False):
    """Worst relative deviation of `state` (the sigma of the benchmark inputs) from the
    signature the UNMODIFIED reference produced in the build container
    (tests/golden/ref_large.npz, tests/golden/make_golden_large.py); None if there is no
    signature for this size."""
    try:
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_golden_large as GL
        with np.load(os.path.join(ROOT, "tests", "golden", "ref_large.npz")) as z:
            tag = f"sigma{norb}_{kind}"
            if f"{tag}_norm" not in z.files:
                return None
            sig = GL.stored_signature(z, tag)
            sig = {k: np.array(v) for k, v in sig.items()}
        seed = int(sig["seed"][0])
        got = GL.signature(state, seed) if numpy_state else GL.signature_torch(state, seed)
        return float(GL.signature_error(sig, got)[0])
    except Exception as exc:  # the checker must not take the measurement down
        sys.stderr.write(f"verify_against_golden: {exc}\n")
        return None


# ----------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------
def measure_fp64_gemm_peak(torch, n=8192, reps=6):
    """cuBLAS DGEMM burst throughput (TFLOP/s): the FP64 roofline denominator."""
    a = torch.randn((n, n), dtype=torch.float64, device="cuda")
    b = torch.randn((n, n), dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n**3 / (best * 1e-3) / 1e12


def run_leg(torch, dist, lib, fqe, args, kind, world, rank, do_e2e, shard=None):
    """Time K sigma builds (device-resident) and, optionally, K end-to-end builds."""
    import ctypes
    from fqe_b200 import synth
    from fqe_b200.distributed import shard_plan, sharded_apply, sharded_apply_host
    from fqe_b200.fqe_data import DenseOperator

    shard = shard or args.shard
    norb = args.norb
    n, sz = norb, 0
    na, nb, la, lb = synth.sector_dims(n, sz, norb)
    h1, h2 = synth.integrals(norb, kind)
    host_c = torch.from_numpy(synth.state(la, lb, seed=synth.seed_for(norb, 50))).pin_memory()
    wfn = fqe.Wavefunction([[n, sz, norb]])
    sector = wfn.sector((n, sz))
    sector.set_wfn(strategy="from_data", raw_data=host_c)
    op = DenseOperator(norb, h1, h2)
    rows, pairs = shard_plan(shard, rank, world, la, op.npair)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return sharded_apply(sector, op, shard)

    for _ in range(args.warmup):
        sigma = step()
    sigma = None
    barrier()
    lib.fqeb_profile_enable(1)
    ms3, cnt3 = (ctypes.c_double * 3)(), (ctypes.c_int64 * 3)()
    lib.fqeb_profile_collect(ms3, cnt3)  # reset
    launches0 = lib.fqeb_launch_count()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        sigma = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.fqeb_launch_count() - launches0
    lib.fqeb_profile_collect(ms3, cnt3)
    lib.fqeb_profile_enable(0)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    path = int(lib.fqeb_sigma_last_path())
    checksum = float(torch.view_as_real(sigma).abs().sum().item())
    # the timed sigma against the reference's own result for these inputs (always on: it is
    # a few reductions on the device, outside the timed region)
    golden_verify = verify_against_golden(sigma, norb, kind)
    verify = None
    if args.verify:
        full = sector.apply_operator(op)
        verify = float((torch.linalg.norm(sigma - full) / torch.linalg.norm(full)).item())
        del full
    del sigma

    result = {
        "kind": kind, "op_kind": op.kind, "op_npair": op.npair, "op_sym": op.symmetric,
        "ms_total": ms_max, "launches": int(launches),
        "phase_ms": [float(x) for x in ms3], "phase_launches": [int(x) for x in cnt3],
        "rows": rows, "pairs": pairs, "la": la, "lb": lb, "clocks": clocks,
        "checksum": checksum, "verify": verify, "golden_verify": golden_verify, "path": path,
    }

    if do_e2e:
        from fqe_b200.distributed import ExchangeBuffers, HostApplyStream, SharedHostBuffer
        # Host buffers.  One rank: pinned tensors.  Several ranks: pinned POSIX shared memory
        # owned by rank 0 and mapped by every rank, so that rank 0's process holds the complete
        # input and receives the complete sigma while every GPU moves only its block of rows.
        shared = []
        if world == 1:
            host_in = host_c
            outs = [torch.empty((la, lb), dtype=torch.complex128).pin_memory() for _ in range(2)]
        else:
            tag = "fqeb_%s_" % os.environ.get("MASTER_PORT", "0")
            if rank == 0:
                shared = [SharedHostBuffer(tag + nm, (la, lb), create=True)
                          for nm in ("c", "s0", "s1")]
                shared[0].tensor.copy_(host_c)
            dist.barrier()
            if rank != 0:
                shared = [SharedHostBuffer(tag + nm, (la, lb)) for nm in ("c", "s0", "s1")]
            host_in = shared[0].tensor
            outs = [shared[1].tensor, shared[2].tensor]
        bufs = ExchangeBuffers(sector, world, rank)

        def e2e_step(k):
            ham = fqe.get_restricted_hamiltonian((h1, h2))
            op_i = wfn._dense_operator(ham.tensors())                # operator preparation
            # H2D of this step's input (each rank its block of rows, completed by one all-gather
            # over NVLink), sharded sigma, one reduce-scatter, D2H of each rank's block
            sharded_apply_host(sector, op_i, host_in, outs[k % 2], shard, bufs)
            torch.cuda.synchronize()

        e2e_step(0)
        barrier()
        if rank == 0:   # the calling process holds the COMPLETE host sigma: check it
            result["e2e_verify"] = verify_against_golden(outs[0].numpy(), norb, kind,
                                                         numpy_state=True)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_step(k)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        result["e2e_serial_s"] = float(t.item())
        del bufs
        # the same K builds as a stream: upload of build k+1 / download of build k-1 overlap
        # build k on separate CUDA streams (every build still copies its own input and result)
        pipe = HostApplyStream(sector, shard)

        def e2e_stream(k):
            ham = fqe.get_restricted_hamiltonian((h1, h2))
            pipe.submit(wfn._dense_operator(ham.tensors()), host_in, outs[k % 2])

        for k in range(2):
            e2e_stream(k)
        pipe.drain()
        barrier()
        if rank == 0:
            result["e2e_stream_verify"] = float(
                (torch.linalg.norm(outs[1] - outs[0]) / torch.linalg.norm(outs[0])).item())
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_stream(k)
        pipe.drain()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        result["e2e_s"] = float(t.item())
        del pipe
        # bytes over PCIe per step, summed over all ranks
        result["h2d"] = host_c.numel() * 16 + world * (h1.nbytes + h2.nbytes)
        result["d2h"] = outs[0].numel() * 16
        # the reference-facing C-ABI call with plain (pageable) numpy buffers, one rank only:
        # fqeb_sigma_restricted_host is what a maintainer binds at src/fqe/fqe_data.py:685
        if world == 1:
            import numpy as np
            from fqe_b200.fqe_data import fold_restricted
            h1p, h2p = fold_restricted(h1, h2)
            c_np = np.array(host_c.numpy(), copy=True)          # pageable
            s_np = np.empty_like(c_np)

            def cabi_step():
                rc = lib.fqeb_sigma_restricted_host(norb, na, nb, h1p.ctypes.data, h2p.ctypes.data,
                                                    c_np.ctypes.data, s_np.ctypes.data)
                if rc != 0:
                    raise RuntimeError(lib.fqeb_last_error().decode())

            cabi_step()
            result["cabi_verify"] = verify_against_golden(s_np, norb, kind, numpy_state=True)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                cabi_step()
            result["cabi_s"] = time.perf_counter() - t0
        barrier()
        for sh in shared:
            sh.close()
    return result


def run_b200(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    import torch.distributed as dist
    import fqe_b200 as fqe
    from fqe_b200 import lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = L.load()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    fp64_peak = measure_fp64_gemm_peak(torch)
    import ctypes as _ct
    _tops = _ct.c_double(0.0)
    i8_peak = float(_tops.value) if lib.fqeb_i8_tensor_peak(_ct.byref(_tops)) == 0 else 0.0

    main = run_leg(torch, dist, lib, fqe, args, args.kind, world, rank, do_e2e=True)
    other = None
    if not args.no_secondary:
        other_kind = "herm" if args.kind == "real8" else "real8"
        other = run_leg(torch, dist, lib, fqe, args, other_kind, world, rank, do_e2e=False)

    pair_leg = None
    if world > 1 and not args.no_secondary and args.shard != "pair":
        # the partition north_star names (pair index sharded), next to the default row sharding
        pair_leg = run_leg(torch, dist, lib, fqe, args, args.kind, world, rank, do_e2e=False,
                           shard="pair")

    def summarise(res):
        npair = res["op_npair"]   # pair space of the contraction (compressed if symmetric)
        la, lb = res["la"], res["lb"]
        r0, r1 = res["rows"]
        p0, p1 = res["pairs"]
        ndet_rank = (r1 - r0) * lb
        cplx = res["op_kind"] == L.OP_COMPLEX
        # algorithmic flops of the contraction on this rank, per sigma
        flop = (8.0 if cplx else 4.0) * npair * (p1 - p0) * ndet_rank
        k = args.steps
        gemm_ms = res["phase_ms"][1]
        achieved = flop * k / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        # algorithmic bytes of gather / scatter: 16*(npair_slice+1) B per det
        gat_b = 16.0 * ((p1 - p0) + 1) * ndet_rank
        sca_b = 16.0 * (npair + 1) * ndet_rank
        phases = {
            "gather": {"ms_per_step": res["phase_ms"][0] / k,
                       "GBps": gat_b * k / (res["phase_ms"][0] * 1e-3) / 1e9
                       if res["phase_ms"][0] > 0 else 0.0},
            "contract": {"ms_per_step": gemm_ms / k, "TFLOPs": achieved},
            "scatter": {"ms_per_step": res["phase_ms"][2] / k,
                        "GBps": sca_b * k / (res["phase_ms"][2] * 1e-3) / 1e9
                        if res["phase_ms"][2] > 0 else 0.0},
        }
        for nm in ("gather", "scatter"):
            phases[nm]["frac_of_hbm_peak"] = phases[nm]["GBps"] / hbm_peak
        share = gemm_ms / max(sum(res["phase_ms"]), 1e-9)
        return flop, achieved, phases, share

    flop, achieved, phases, share = summarise(main)
    value = args.steps / (main["ms_total"] * 1e-3)
    # DRAM traffic of the dominant kernel per launch, from the committed ncu capture of the
    # same workload (profiles/r01_traffic.json); only quoted for the configuration it was
    # taken on (norb=16, real pair-symmetric operator, one GPU)
    traffic = None
    fused = main["phase_launches"][0] == 0 and main["phase_launches"][1] > 0
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if args.norb == 16 and world == 1 and main["op_sym"] and main["op_kind"] != L.OP_COMPLEX:
            traffic = tr["fused"]["dram_bytes_per_launch"] if fused else tr["dram_bytes_per_launch"]
    except Exception:
        pass
    traffic_sliced = None
    try:
        tr2 = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if args.norb == 16 and world == 1 and main["op_sym"] and main["op_kind"] != L.OP_COMPLEX:
            traffic_sliced = tr2["sliced"]["dram_bytes_per_launch"]
    except Exception:
        pass
    gemm_s = main["phase_ms"][1] * 1e-3
    launches_per_step = main["phase_launches"][1] // max(args.steps, 1)
    flop_model = "%d*P^2*L^2 with P=%d pairs (%s; %s)" % (
        8 if main["op_kind"] == L.OP_COMPLEX else 4, main["op_npair"],
        "complex h2'" if main["op_kind"] == L.OP_COMPLEX else
        "real h2' times complex D: 2 real FMAs per element",
        "i>=j compressed pair space, h2' pair-symmetric" if main["op_sym"] else
        "full norb^2 pair space")
    if main["path"] == 3:
        # INT8-sliced contraction: the tensor work actually executed is 21 exact INT8 slice
        # products of the same P x P by P x 2*ndet shape; its roofline is the INT8 tensor rate
        # measured in this run.  The FP64 figures say what that buys on the algorithmic flops.
        r0, r1 = main["rows"]
        ndet_rank = (r1 - r0) * main["lb"]
        nprod = 21
        i8_ops = 2.0 * nprod * main["op_npair"] ** 2 * 2.0 * ndet_rank       # per sigma, this rank
        ach_i8 = i8_ops * args.steps / gemm_s / 1e12 if gemm_s > 0 else 0.0
        roofline = {
            "bound": "tensor",
            "kernel": "k_sigma_ozaki2 (gather of digit planes into the UMMA tile + 21 INT8 slice "
                      "products per tile on tcgen05.mma kind::i8, INT32 accumulators in TMEM, "
                      "FP64 reconstruction in the epilogue; the gather phase is inside this kernel)",
            "achieved": ach_i8, "peak": i8_peak, "unit": "TOP/s",
            "frac": ach_i8 / i8_peak if i8_peak > 0 else None, "traffic": traffic_sliced,
            "traffic_unit": "DRAM bytes per launch of the sliced contraction (ncu), see "
                            "profiles/r02_traffic.json; quoted only for the configuration it was "
                            "captured on (norb=16, one GPU)",
            "peak_source": "tcgen05.mma kind::i8 M=128 N=256 K=32 on all SMs, CUDA-event timed, "
                           "this run (fqeb_i8_tensor_peak); MEASURED_PEAKS.json has bf16 only",
            "ops_model": "2 * 21 slice products * P^2 * 2*ndet INT8 multiply-adds with P=%d "
                         "(6 radix-127 digit slices per operand, products i+j<=5 kept)" %
                         main["op_npair"],
            "ops_per_step_per_rank": i8_ops,
            "launches_per_step": launches_per_step,
            "fp64_equivalent": {
                "achieved": achieved, "unit": "TFLOP/s", "flops_per_step_per_rank": flop,
                "flop_model": flop_model, "fp64_gemm_peak": fp64_peak,
                "ratio_to_fp64_gemm_peak": achieved / fp64_peak if fp64_peak > 0 else None,
                "fp64_peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul(float64), best of 6, "
                                    "this run"},
            "share_of_step": share,
        }
    else:
        roofline = {
            "bound": "tensor",
            "kernel": ("k_sigma_fused (D tiles gathered into the shared-memory ring + FP64 DMMA "
                       "contraction; the gather phase is inside this kernel)") if fused else
                      "k_dgemm_ws (FP64 DMMA contraction)",
            "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
            "traffic_unit": "DRAM bytes per contraction launch (ncu), see profiles/r01_traffic.json",
            "launches_per_step": launches_per_step,
            "flops_per_step_per_rank": flop,
            "flop_model": flop_model,
            "peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul(float64), best of 6, this run",
            "share_of_step": share,
        }
    line = {
        "metric": metric_name(args.norb), "value": value, "unit": "sigma/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_total"] / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(
            args.norb, args.kind, args.shard, world,
            {L.OP_REAL: "real", L.OP_IMAG: "imag", L.OP_COMPLEX: "complex"}[main["op_kind"]] +
            (", pair-symmetric" if main["op_sym"] else "")),
        "gpu_launches": main["launches"],
        "verify_rel_err": main["golden_verify"], "verify_against": "signature of the UNMODIFIED "
        "reference's sigma for these inputs (tests/golden/ref_large.npz)",
        "shard_verify_rel_err": main["verify"], "e2e_verify_rel_err": main.get("e2e_verify"),
        "contraction_path": {0: "none", 1: "gather -> FP64 DMMA GEMM -> scatter",
                             2: "gather fused into the FP64 DMMA contraction",
                             3: "gather fused into the INT8-sliced tcgen05 contraction"}.get(
            main.get("path"), None),
        "clocks": main["clocks"],
        "e2e": {"value": args.steps / main["e2e_serial_s"], "unit": "sigma/s",
                "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": main["d2h"],
                "mode": "K builds strictly one after the other through the public host-buffer API "
                        "(fqe_b200.distributed.sharded_apply_host + synchronize per build): "
                        "operator preparation, upload from pinned host memory, sigma, download; "
                        "with N > 1 ranks the host arrays are pinned POSIX shared memory owned by "
                        "rank 0, every rank moves its block of rows (one all-gather before, one "
                        "reduce-scatter after the build) and rank 0 holds the complete host sigma",
                "verify_rel_err": main.get("e2e_verify"),
                "streamed_value": args.steps / main["e2e_s"],
                "streamed_mode": "the same K builds software-pipelined (HostApplyStream): upload "
                                 "of build k+1 and download of build k-1 overlap build k on "
                                 "separate CUDA streams",
                "stream_verify_rel_err": main.get("e2e_stream_verify"),
                "cabi_host_value": (args.steps / main["cabi_s"]) if main.get("cabi_s") else None,
                "cabi_host_mode": "fqeb_sigma_restricted_host through ctypes with plain pageable "
                                  "numpy arrays in and out (the call INTEGRATION.md binds at "
                                  "src/fqe/fqe_data.py:685), one GPU",
                "cabi_host_verify_rel_err": main.get("cabi_verify")},
        "roofline": roofline,
        "phases": phases,
        "hbm_peak": {"GBps": hbm_peak, "source": hbm_src},
    }
    if pair_leg is not None:
        line["pair_shard"] = {
            "value": args.steps / (pair_leg["ms_total"] * 1e-3), "unit": "sigma/s",
            "ms_per_step": pair_leg["ms_total"] / args.steps,
            "verify_rel_err": pair_leg["golden_verify"],
            "note": "dvec pair index ij sharded over the ranks (BASELINE.json north_star), C "
                    "replicated, one all-reduce; gather and contraction shrink with the world "
                    "size, every rank still scatters all pair rows of its partial E",
        }
    if other is not None:
        f2, a2, ph2, sh2 = summarise(other)
        line["secondary"] = {
            "workload": workload_name(args.norb, other["kind"]),
            "value": args.steps / (other["ms_total"] * 1e-3), "unit": "sigma/s",
            "roofline": {"achieved": a2, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": a2 / fp64_peak if fp64_peak > 0 else None,
                         "flops_per_step_per_rank": f2, "share_of_step": sh2},
            "phases": ph2,
        }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            threads = use_all_host_threads()
            secs, desc = cpu_sigma_seconds(args.norb, args.kind, args.cpu_budget)
            line["cpu_baseline"] = {"value": 1.0 / secs, "unit": "sigma/s", "cores": host_cores(),
                                    "omp_threads": threads, "kind": "reference",
                                    "extrapolated": True,
                                    "sample": desc + "; the full-size measurement is the "
                                    "`--impl reference` arm (this bounded sample under-states the "
                                    "reference's time at full size, see DESIGN.md 7)"}
        except Exception as exc:  # the checker is optional for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "sigma/s", "cores": host_cores(),
                                    "kind": "reference", "sample": f"unavailable: {exc}"}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def quiet_stdout():
    """Send everything libraries write to stdout (NCCL prints its version banner there) to
    stderr, and keep the original stdout for the ONE JSON line of the contract."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
