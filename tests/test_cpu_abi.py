"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a,
loads without a GPU, exports every symbol include/ppb200.h declares, its
structs have the layout the ctypes binding assumes, and compute calls fail
loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ppb200.h")


@pytest.fixture(scope="module")
def ffi():
    from pulseportraiture_b200 import _ffi
    _ffi.build()
    return _ffi


def header_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(ffi):
    L = ffi.lib()
    declared = header_functions()
    assert declared, "no functions parsed from the header"
    for name in declared:
        assert hasattr(L, name), "missing symbol " + name
    assert sorted(ffi.SYMBOLS) == declared
    assert L.pp_abi_version() == 6


def test_struct_layouts_match_the_header(ffi, tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ppb200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(pp_fit_args_t), offsetof(pp_fit_args_t, nu_fit_mode),'
                   'offsetof(pp_fit_args_t, fit_flags), offsetof(pp_fit_args_t, tol),'
                   'sizeof(pp_fit_out_t), sizeof(pp_pshift_out_t), sizeof(pp_stats_t),'
                   'offsetof(pp_stats_t, chunk));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(ffi.FitArgs), ffi.FitArgs.nu_fit_mode.offset, ffi.FitArgs.fit_flags.offset,
            ffi.FitArgs.tol.offset, C.sizeof(ffi.FitOut), C.sizeof(ffi.PShiftOut),
            C.sizeof(ffi.Stats), ffi.Stats.chunk.offset]
    assert got == want


def test_no_cpu_fallback(ffi):
    """Without a CUDA device plan creation must fail with a message; with one
    (GPU box) it must succeed.  Either way nothing is computed on the CPU."""
    L = ffi.lib()
    h = C.c_void_p()
    rc = L.pp_plan_create(8, 256, 0, C.byref(h))
    if rc == 0:
        L.pp_plan_destroy(h)
    else:
        assert rc < 0 and len(L.pp_last_error()) > 0
    assert L.pp_plan_create(8, 1001, 0, C.byref(h)) < 0        # odd nbin: no real-FFT packing
    assert b"even" in L.pp_last_error()
    assert L.pp_plan_create(8, 8192, 0, C.byref(h)) < 0 and b"4096" in L.pp_last_error()
    assert L.pp_fit_batch(None, None, None) < 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pulseportraiture_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("pp_oracle-free", ""), f


def test_host_helpers_match_oracle():
    """Host-side (non-hot-path) helpers of the facade: model generator and
    scalar transforms agree with the oracle's independent restatement."""
    from pulseportraiture_b200 import pplib
    from oracle import pp_oracle as orc
    from tests import synth
    freqs, model = synth.example_model(16, 256, 1500., 800.)
    _, _, m2 = pplib.read_model(synth.GMODEL, pplib.get_bin_centers(256), freqs,
                                synth.P_EXAMPLE, quiet=True)
    assert np.allclose(m2, model, rtol=0, atol=1e-12)
    assert pplib.guess_fit_freq(freqs) == pytest.approx(orc.guess_fit_freq(freqs), rel=1e-15)
    for a in ((0.3, 2e-3, 1400., 1500.), (0.49, 5e-3, 1200., np.inf)):
        assert pplib.phase_transform(*a, P=synth.P_EXAMPLE, mod=True) == \
            pytest.approx(orc.phase_transform(*a, P=synth.P_EXAMPLE, mod=True), abs=1e-15)
    name, code, nu_ref, ngauss, params, flags, alpha, fit_alpha = pplib.read_model(synth.GMODEL)
    assert ngauss == 3 and code == "000" and nu_ref == 1300.0 and alpha == -4.0


def test_mjd_arithmetic():
    from pulseportraiture_b200.pptoas import MJD
    t = MJD(55000, 0.75) + MJD(0, 0.5)
    assert t.intday() == 55001 and abs(t.fracday() - 0.25) < 1e-15
    assert abs((MJD(55000, 0.1) + 1e-9).in_days() - 55000.100000001) < 1e-9


def test_write_TOAs_matches_reference_lines(tmp_path):
    """pplib.write_TOAs / filter_TOAs / TOA.write_TOA against lines written by the reference's own
    write_TOAs (tests/golden/make_golden_toas.py -> toas_v1.tim)."""
    import json
    import os
    import numpy as np
    from pulseportraiture_b200 import pplib, pptoas
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    specs = json.load(open(os.path.join(gdir, "toas_v1.json")))
    want = open(os.path.join(gdir, "toas_v1.tim")).read()
    toas = [pptoas.TOA(s["archive"], np.inf if s["frequency"] == "inf" else s["frequency"],
                       pptoas.MJD(s["mjd"][0], s["mjd"][1]), s["err"], s["telescope"], s["code"], s["DM"],
                       s["DM_error"], dict(s["flags"])) for s in specs]
    out = str(tmp_path / "out.tim")
    pplib.write_TOAs(toas, inf_is_zero=True, SNR_cutoff=8.0, outfile=out, append=False)
    pplib.write_TOAs([toas[1]], inf_is_zero=False, SNR_cutoff=0.0, outfile=out, append=True)
    assert open(out).read() == want
    kept, culled = pplib.filter_TOAs(toas, "snr", 8.0, ">=", return_culled=True)
    assert [t.archive for t in kept] == [specs[0]["archive"], specs[1]["archive"]] and len(culled) == 2
    one = str(tmp_path / "one.tim")
    toas[0].write_TOA(outfile=one)
    assert open(one).read() == want.splitlines()[0] + "\n"


def test_plain_c_client_compiles_and_links(tmp_path):
    """tests/c/abi_fit.c (C99, -Wall -Werror) builds against include/ppb200.h and links with the library:
    the header is usable from C and every symbol the client needs is exported (run: test_gpu_c_client)."""
    import os
    import subprocess
    from pulseportraiture_b200 import _ffi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _ffi.lib()                                   # builds the library if needed
    libdir = os.path.dirname(_ffi.LIB_PATH)
    exe = str(tmp_path / "abi_fit")
    r = subprocess.run(["gcc", "-O1", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "c", "abi_fit.c"), "-o", exe, "-L", libdir, "-lppb200",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.isfile(exe)


def test_bounds_and_frequency_table_helpers():
    """Host-side argument handling added with the bounds / per-subint frequency tables (no GPU)."""
    from pulseportraiture_b200 import pplib, pptoas
    from pulseportraiture_b200.engine import bounds_array
    # scipy-style bounds -> [5, 2] with NaN for open ends; nothing bounded -> None (NULL in the ABI)
    assert bounds_array(None) is None
    assert bounds_array([(None, None)] * 5) is None
    b = bounds_array([(None, None), (None, 1e-3), None, (-2.5, None), (-10.0, 10.0)])
    assert b.shape == (5, 2) and b.dtype == np.float64
    assert np.isnan(b[0]).all() and np.isnan(b[1, 0]) and b[1, 1] == 1e-3 and np.isnan(b[2]).all()
    assert b[3, 0] == -2.5 and np.isnan(b[3, 1]) and list(b[4]) == [-10.0, 10.0]
    assert pplib._check_bounds([(None, None), (None, None)], 2) is None
    assert pplib._check_bounds([(0.0, None), (None, None)], 2) == [(0.0, None), (None, None)]
    with pytest.raises(ValueError):
        pplib._check_bounds([(None, None)] * 3, 2)
    with pytest.raises(ValueError):
        pplib._check_bounds([(0.0, 1.0, 2.0)], 2)
    # distinct frequency tables of an archive and the table each subint uses
    fA, fB = np.linspace(1100., 1900., 8), np.linspace(1100.5, 1900.5, 8)
    tables, table_of = pptoas._freq_tables(np.stack([fA, fB, fA, fA, fB]))
    assert tables.shape == (2, 8) and list(table_of) == [0, 1, 0, 0, 1]
    assert np.array_equal(tables[0], fA) and np.array_equal(tables[1], fB)
    tables, table_of = pptoas._freq_tables(np.tile(fA, (4, 1)))
    assert tables.shape == (1, 8) and list(table_of) == [0, 0, 0, 0]


def test_round2_host_helpers():
    """Host-side helpers added in round 2 (no GPU): the batched guess_fit_freq, the model-array policy of the
    facade, the scipy return-code mapping."""
    from pulseportraiture_b200 import pplib
    from oracle import pp_oracle as orc
    rng = np.random.RandomState(11)
    f = np.sort(rng.uniform(400.0, 800.0, (6, 40)), axis=1)
    snr = rng.uniform(0.0, 5.0, (6, 40))
    m = rng.rand(6, 40) > 0.3
    m[4] = False                                             # a subint without usable channels
    got = pplib.guess_fit_freq_batch(f, snr, m)
    for i in range(6):
        if m[i].any():
            assert abs(got[i] - pplib.guess_fit_freq(f[i, m[i]], snr[i, m[i]])) < 1e-10
            assert abs(got[i] - orc.guess_fit_freq(f[i, m[i]], snr[i, m[i]])) < 1e-10
        else:
            assert got[i] == 0.0
    # models: float32 arrays stay float32, everything else reaches the device as float64 (pp_set_model_f64)
    a64 = np.arange(12.0).reshape(3, 4)
    assert pplib._mdl(a64).dtype == np.float64 and pplib._mdl(a64.astype(np.float32)).dtype == np.float32
    assert pplib._mdl(a64.astype(np.int32)).dtype == np.float64 and pplib._mdl(a64[:, ::2]).flags.c_contiguous
    # device solver codes -> the codes the reference's scipy method would report for the same outcome
    assert [pplib.scipy_return_code(c, "TNC") for c in (0, 1, 3)] == [1, 3, 6]
    assert [pplib.scipy_return_code(c, "trust-ncg") for c in (0, 1, 3)] == [2, 1, 3]
    assert pplib.scipy_return_code(0, "Newton-CG") == 0
