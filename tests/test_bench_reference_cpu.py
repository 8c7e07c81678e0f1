"""The reference arm of bench.py (`--impl reference`) on CPU at a small size: it must run the
reference's C kernels on SEVERAL cores.  (Round 2 found it pinned to one thread by an
MKL_NUM_THREADS setting while omp_get_max_threads() reported all of them; the line now carries
`busy_cores` = process CPU time / wall time of the timed builds.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref")),
                    reason="oracle/_ref not built")
def test_reference_arm_uses_the_host_cores():
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")   # what torchrun exports
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--norb", "12", "--steps", "3", "--warmup", "0", "--cpu-budget", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "sigma/s" and line["steps"] >= 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["extrapolated"] is False
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cores = cb["cores"]
    assert cb["omp_threads"] == cores
    if cores >= 4:
        assert cb["busy_cores"] > 2.0, cb
        assert "warning" not in cb
