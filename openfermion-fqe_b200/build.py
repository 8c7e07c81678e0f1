#!/usr/bin/env python
"""Build libfqe_b200.so in-tree with nvcc for sm_100a.

    python openfermion-fqe_b200/build.py [--force] [--verbose] [--define NAME ... --out PATH]

``--define`` / ``--out`` build an experimental variant next to the product library (A/B runs
through the ``FQEB_B200_LIB`` environment variable, see fqe_b200/lib/__init__.py).

The library is a plain C-ABI shared object (see include/fqe_b200.h); it does not
link against torch or Python.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "fqe_b200", "lib")
OUT = os.path.join(OUT_DIR, "libfqe_b200.so")
SOURCES = ["core.cu", "graph.cu", "blas1.cu", "dcoulomb.cu", "dvec.cu", "dgemm.cu", "sigma.cu",
           "rotate.cu", "nbody.cu", "ozaki.cu", "rdm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
    "-ccbin", "/usr/bin/g++", "-Xptxas=-v", "-cudart", "static",
]


HEADERS = [os.path.join(ROOT, "include", "fqe_b200.h")]
OBJ_DIR = os.path.join(HERE, "build", "obj")


def _headers():
    return HEADERS + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
        [os.path.abspath(__file__)]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    src, obj, defines, verbose = args
    cmd = [NVCC] + [f for f in FLAGS if f != "-shared"] + ["-D" + d for d in defines] + \
        ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res


def build(force=False, verbose=False, defines=(), out=None):
    """One object per source file (compiled in parallel, rebuilt only when the source or a
    header changed), linked into the shared library."""
    if out is None and not force and not stale():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OUT_DIR, exist_ok=True)
    variant = "default" if not defines else "_".join(sorted(defines)).replace("=", "-")
    obj_dir = os.path.join(OBJ_DIR, variant)
    os.makedirs(obj_dir, exist_ok=True)
    out = out or OUT
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    jobs, objs = [], []
    for name in SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(obj_dir, name.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or \
                os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj, tuple(defines), verbose))
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as pool:
        for src, res in pool.map(_compile_one, jobs):
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed compiling {os.path.basename(src)}")
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin",
            "/usr/bin/g++", "-cudart", "static"] + objs + ["-o", out]
    res = subprocess.run(link, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed linking libfqe_b200.so")
    return out


if __name__ == "__main__":
    argv = sys.argv[1:]
    defs = [argv[i + 1] for i, a in enumerate(argv) if a == "--define"]
    outp = argv[argv.index("--out") + 1] if "--out" in argv else None
    print(build(force="--force" in argv, verbose="--verbose" in argv or "-v" in argv,
                defines=defs, out=outp))
