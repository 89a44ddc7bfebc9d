"""The C ABI used from plain C: tests/c/abi_fit.c is compiled with gcc against include/ppb200.h and
libppb200.so and must give the numbers the Python wrapper gives (bit for bit: same library)."""
import os
import subprocess

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_client_matches_python_wrapper(tmp_path):
    from pulseportraiture_b200 import _ffi
    from pulseportraiture_b200.engine import WidebandPlan
    libdir = os.path.dirname(_ffi.LIB_PATH)
    exe = str(tmp_path / "abi_fit")
    cmd = ["gcc", "-O1", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_fit.c"), "-o", exe, "-L", libdir, "-lppb200",
           "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    nsub, nchan, nbin = 3, 32, 512
    cases = [synth.make_case(nchan, nbin, 1500., 800., 8100 + s) for s in range(nsub)]
    data = np.stack([c["data"] for c in cases]).astype(np.float32)
    model, freqs, P = cases[0]["model"].astype(np.float32), cases[0]["freqs"], cases[0]["P"]
    raw = str(tmp_path / "in.bin")
    with open(raw, "wb") as fh:
        fh.write(np.asarray(freqs, dtype=np.float64).tobytes())
        fh.write(model.tobytes())
        fh.write(data.tobytes())
    run = subprocess.run([exe, raw, str(nchan), str(nbin), str(nsub), repr(float(P))], capture_output=True,
                         text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    rows = np.array([[float(x) for x in ln.split()] for ln in run.stdout.strip().splitlines()])
    with WidebandPlan(nchan, nbin) as pl:
        pl.set_model(model, freqs)
        r = pl.fit_batch(data, P)
    assert np.array_equal(rows[:, 1], r["params"][:, 0]) and np.array_equal(rows[:, 3], r["params"][:, 1])
    assert np.array_equal(rows[:, 2], r["param_errs"][:, 0]) and np.array_equal(rows[:, 5], r["chi2"])
    assert np.array_equal(rows[:, 6], r["nu_out"][:, 0])
    assert np.array_equal(rows[:, 7].astype(int), r["return_code"]) and np.array_equal(rows[:, 8].astype(int), r["lag_index"])
