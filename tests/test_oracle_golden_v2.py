"""The CPU oracle against golden_v2.npz: outputs of the reference's own functions at the BASELINE.json
shapes (512 x 2048 phi+DM; 256 / 512 x 1024 five-parameter fits at sigma = 1.5), every fit_flags pattern
again at sigma = 1.5, the instrumental response and the noise variants (tests/golden/make_golden_v2.py).
Same bars as test_oracle_golden.py: parameters 1e-4 sigma, chi2 1e-10."""
import os

import numpy as np
import pytest

from oracle import pp_oracle as orc
from tests import synth
from tests.test_oracle_golden import rel

G2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v2.npz"))


def cases(prefix):
    return sorted({k.split("/")[0] for k in G2.files if k.startswith(prefix)})


def check_fp(case, tag, r, sig_tol=1e-4):
    g = lambda f: G2["%s/%s.%s" % (case, tag, f)]  # noqa: E731
    assert abs(r.phase - g("phase")) / g("phase_err") < sig_tol
    assert abs(r.DM - g("DM")) / g("DM_err") < sig_tol
    assert rel(r.phase_err, g("phase_err")) < 1e-6
    assert rel(r.DM_err, g("DM_err")) < 1e-6
    assert rel(r.nu_ref, g("nu_ref")) < 1e-6
    assert rel(r.chi2, g("chi2")) < 1e-10
    assert rel(r.snr, g("snr")) < 1e-8
    assert rel(r.scales, g("scales")) < 1e-5


def check_full(case, tag, r, flags, sig_tol=1e-4, nu_tol=1e-6):
    g = lambda f: G2["%s/%s.%s" % (case, tag, f)]  # noqa: E731
    for i, nm in enumerate(["phi", "DM", "GM", "tau", "alpha"]):
        if flags[i]:
            assert abs(r[nm] - g(nm)) / g(nm + "_err") < sig_tol, nm
            assert rel(r[nm + "_err"], g(nm + "_err")) < 1e-5, nm
        else:
            assert abs(r[nm] - g(nm)) <= 1e-12 * max(1.0, abs(g(nm))), nm
    for nm in ("nu_DM", "nu_GM", "nu_tau"):
        assert rel(r[nm], g(nm)) < nu_tol, nm
    assert rel(r.chi2, g("chi2")) < 1e-10
    assert rel(r.snr, g("snr")) < 1e-8
    assert rel(r.scales, g("scales")) < 1e-5
    assert rel(r.scale_errs, g("scale_errs")) < 1e-5


@pytest.mark.parametrize("case", cases("b2_"))
def test_config2_shape_phidm(case):
    nchan, nbin, nu0, bw, seed = G2[case + "/cfg"]
    c = synth.make_case(int(nchan), int(nbin), nu0, bw, int(seed))
    assert np.allclose(synth.checksum(c["data"]), G2[case + "/in_checksum"], rtol=1e-13)
    data, model, freqs, P = c["data"], c["model"], c["freqs"], c["P"]
    errs = orc.get_noise(data, chans=True)
    assert rel(errs, G2[case + "/noise"]) < 1e-12
    g = orc.fit_phase_shift(data.mean(0), model.mean(0), Ns=100)
    assert abs(g.phase - G2[case + "/ps.phase"]) < 1e-8
    check_fp(case, "fp", orc.fit_portrait(data, model, np.array([g.phase, 0.0]), P, freqs, errs=errs))
    r = orc.fit_portrait_full(data, model, [g.phase, 0.0, 0.0, 0.0, 0.0], P, freqs, errs=errs,
                              fit_flags=[1, 1, 0, 0, 0], log10_tau=False)
    check_full(case, "full", r, [1, 1, 0, 0, 0])


def run_full_case(case):
    cfg = G2[case + "/cfg"]
    nchan, nbin, nu0, bw, seed = int(cfg[0]), int(cfg[1]), cfg[2], cfg[3], int(cfg[4])
    tau_s, log10, option, sigma = cfg[5], bool(cfg[6]), int(cfg[7]), cfg[8]
    flags = [int(v) for v in G2[case + "/flags"]]
    c = synth.make_case(nchan, nbin, nu0, bw, seed, tau_data_s=tau_s, sigma=sigma)
    assert np.allclose(synth.checksum(c["data"]), G2[case + "/in_checksum"], rtol=1e-13)
    r = orc.fit_portrait_full(c["data"], c["model"], list(G2[case + "/init"]), c["P"], c["freqs"],
                              errs=G2[case + "/errs"], fit_flags=flags, log10_tau=log10, option=option)
    check_full(case, "full", r, flags)


@pytest.mark.parametrize("case", cases("b3_"))
def test_config3_parity_shapes(case):
    run_full_case(case)


@pytest.mark.parametrize("case", cases("full15_"))
def test_every_flag_pattern_at_sigma_1p5(case):
    run_full_case(case)


def test_instrumental_response():
    nbin = 256
    assert rel(orc.instrumental_response_FT(nbin, 0.013, 'rect'), G2["ir/rect"]) < 1e-13
    gs = orc.instrumental_response_FT(nbin, 0.02, 'gauss')
    assert np.abs(gs.imag).max() <= 1e-15 and float(G2["ir/gauss_imag_max"]) <= 1e-15     # real for loc = 0
    assert np.allclose(gs.real, G2["ir/gauss"], rtol=1e-12, atol=1e-300)
    # (the reference's wid = 0 branch falls off the end and returns None, pptoaslib.py:131-132: not pinned)
    assert np.allclose(orc.gaussian_profile_FT(nbin, 0.3, 0.05, 2.0), G2["ir/gprof_FT"], rtol=1e-12, atol=1e-300)
    freqs, model = synth.example_model(32, nbin, 1500., 800.)
    model = model.astype(np.float32).astype(np.float64)
    for tag in ("wids", "dm", "both"):
        DM, wids, types = float(G2["ir_%s/DM" % tag]), list(G2["ir_%s/wids" % tag]), [str(t) for t in G2["ir_%s/types" % tag]]
        resp = orc.instrumental_response_port_FT(nbin, freqs, DM, synth.P_EXAMPLE, wids, types)
        assert np.allclose(resp.real, G2["ir_%s/resp_real" % tag], rtol=1e-12, atol=1e-300)
        assert float(G2["ir_%s/resp_imag_max" % tag]) < 1e-15
        conv = orc.add_instrumental_response(model, freqs, DM, synth.P_EXAMPLE, wids, types)
        assert np.allclose(conv, G2["ir_%s/conv" % tag], rtol=0, atol=1e-12 * np.abs(conv).max())
    okc = G2["ir_subset/okc"]
    resp = orc.instrumental_response_port_FT(nbin, freqs[okc], 30.0, synth.P_EXAMPLE, [], [])
    assert np.allclose(resp.real, G2["ir_subset/resp_real"], rtol=1e-12, atol=1e-300)


def test_noise_variants():
    c = synth.make_case(16, 512, 1500., 800., 701)
    data = c["data"]
    assert np.allclose(synth.checksum(data), G2["noise/in_checksum"], rtol=1e-13)
    for frac in (1, 2, 4, 8):
        assert rel(orc.get_noise_PS(data, frac=frac, chans=True), G2["noise/ps_frac%d" % frac]) < 1e-12
    assert rel(orc.get_noise_PS(data[3], frac=8), G2["noise/ps_prof_frac8"]) < 1e-12
    assert rel(orc.get_noise_fit(data, chans=True), G2["noise/fit_chans"]) < 1e-12
    assert rel(orc.get_noise_fit(data[3]), G2["noise/fit_prof"]) < 1e-12
    assert rel(orc.get_noise_fit(data, fact=2.0, chans=True), G2["noise/fit_fact2"]) < 1e-12


@pytest.mark.parametrize("case", cases("nb_"))
def test_non_power_of_two_nbin(case):
    nchan, nbin, nu0, bw, seed = G2[case + "/cfg"]
    nchan, nbin, seed = int(nchan), int(nbin), int(seed)
    c = synth.make_case(nchan, nbin, nu0, bw, seed)
    assert np.allclose(synth.checksum(c["data"]), G2[case + "/in_checksum"], rtol=1e-13)
    data, model, freqs, P = c["data"], c["model"], c["freqs"], c["P"]
    errs = orc.get_noise(data, chans=True)
    assert rel(errs, G2[case + "/noise"]) < 1e-12
    assert np.allclose(orc.rotate_data(data, 0.05, 1e-3, P, freqs, 1400.0)[1], G2[case + "/rot_row1"], atol=1e-10)
    lag, grid, vals = orc.fit_phase_shift_grid(data.mean(0), model.mean(0), Ns=100)
    assert lag == int(G2[case + "/lag"]) and rel(vals, G2[case + "/grid_vals"]) < 1e-9
    g = orc.fit_phase_shift(data.mean(0), model.mean(0), Ns=100)
    assert abs(g.phase - G2[case + "/ps.phase"]) < 1e-8
    check_fp(case, "fp", orc.fit_portrait(data, model, np.array([g.phase, 0.0]), P, freqs, errs=errs))
    r = orc.fit_portrait_full(data, model, [g.phase, 0.0, 0.0, 0.0, 0.0], P, freqs, errs=errs,
                              fit_flags=[1, 1, 0, 0, 0], log10_tau=False)
    check_full(case, "full", r, [1, 1, 0, 0, 0])


@pytest.mark.parametrize("case", cases("nbfull_"))
def test_non_power_of_two_nbin_five_parameters(case):
    run_full_case(case)
