"""Wavefunction files: readable from / by the reference (reference wavefunction.py:726-765).

tests/golden/ref_wfn_save.bin was written by the UNMODIFIED reference's Wavefunction.save
(tests/golden/make_golden.py --only wfnio) together with the coefficients it held
(ref_wfn_save.npz).  No GPU needed: fqe_b200.wfn_io is host-only."""
import io
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from fqe_b200 import wfn_io


def test_reads_a_file_written_by_the_reference(golden_dir):
    want = np.load(os.path.join(golden_dir, "ref_wfn_save.npz"))
    with open(os.path.join(golden_dir, "ref_wfn_save.bin"), "rb") as fh:
        data = wfn_io.load(fh)
    assert data["norb"] == int(want["norb"][0])
    assert data["conserved"] == {"n": int(want["n"][0]), "s_z": int(want["sz"][0])}
    assert data["conserve_spin"] and data["conserve_number"]
    key = (int(want["n"][0]), int(want["sz"][0]))
    assert list(data["sectors"]) == [key]
    assert np.array_equal(data["sectors"][key], want["coeff"])


def test_round_trip_and_old_layout():
    rng = np.random.default_rng(3)
    c = rng.standard_normal((6, 4)) + 1j * rng.standard_normal((6, 4))
    buf = io.BytesIO()
    wfn_io.dump(buf, {"n": 3, "s_z": 1}, 4, {(3, 1): (2, 1, c)})
    buf.seek(0)
    data = wfn_io.load(buf)
    assert data["norb"] == 4 and data["conserved"] == {"n": 3, "s_z": 1}
    assert np.array_equal(data["sectors"][(3, 1)], c)
    assert "fqe" not in sys.modules and "fqe.fqe_data" not in sys.modules   # placeholders removed
    # the round-1 layout of this package stays readable
    old = io.BytesIO(pickle.dumps([{"n": 3, "s_z": 1}, 4, [(3, 1), c]]))
    assert np.array_equal(wfn_io.load(old)["sectors"][(3, 1)], c)


def test_refuses_foreign_classes():
    evil = io.BytesIO(pickle.dumps([{}, {}, True, True, 2, [(1, 1), subprocess.Popen]]))
    with pytest.raises(pickle.UnpicklingError):
        wfn_io.load(evil)
    with pytest.raises(ValueError):
        wfn_io.load(io.BytesIO(pickle.dumps({"not": "a list"})))


@pytest.mark.skipif(not os.path.isdir("/tmp/fqe_ref_build/src/fqe"),
                    reason="needs the scratch build of the reference (build container only)")
def test_reference_reads_a_file_written_here(tmp_path, golden_dir):
    rng = np.random.default_rng(5)
    c = rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6))
    with open(tmp_path / "ours.bin", "wb") as fh:
        wfn_io.dump(fh, {"n": 4, "s_z": 0}, 4, {(4, 0): (2, 2, c)})
    np.save(tmp_path / "c.npy", c)
    script = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {golden_dir!r})
        import make_golden as MG
        MG.install_stubs()
        sys.path.insert(0, "/tmp/fqe_ref_build/src")
        import fqe
        w = fqe.Wavefunction([[4, 0, 4]])
        w.read("ours.bin", {str(tmp_path)!r})
        assert w.norb() == 4 and list(w.sectors()) == [(4, 0)]
        assert type(w.sector((4, 0))).__module__ == "fqe.fqe_data"
        assert np.array_equal(w.get_coeff((4, 0)), np.load({str(tmp_path / 'c.npy')!r}))
        out = w.apply((np.eye(4, dtype=complex),))          # the object is fully functional
        assert np.allclose(out.get_coeff((4, 0)), 4 * w.get_coeff((4, 0)))
        print("ok")
    """)
    res = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]
