"""GPU parity at the BASELINE.json shapes against outputs of the REFERENCE's own functions
(tests/golden/golden_v2.npz, made by tests/golden/make_golden_v2.py), through the C ABI, at the bars
north_star states: parameters within 1e-3 sigma, chi2 within 1e-8 RELATIVE (plain, no slack), on
  b2_*      512 x 2048 phi+DM (config 2): reference fit_portrait (TNC) and fit_portrait_full (trust-ncg)
  b3_*      256 x 1024 and 512 x 1024 five-parameter fits [1,1,0,1,1] / [1,1,1,1,1] (config 3 parity shapes)
  full15_*  every fit_flags pattern get_nu_zeros distinguishes, sigma = 1.5
plus the instrumental response and the noise-estimate variants."""
import os

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu

SIG_TOL = 1e-3
CHI2_TOL = 1e-8
G2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v2.npz"))


def cases(prefix):
    return sorted({k.split("/")[0] for k in G2.files if k.startswith(prefix)})


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("case", cases("b2_"))
def test_config2_shape_against_reference(case):
    from pulseportraiture_b200 import pplib, pptoaslib
    nchan, nbin, nu0, bw, seed = G2[case + "/cfg"]
    c = synth.make_case(int(nchan), int(nbin), nu0, bw, int(seed))
    data, model, freqs, P = c["data"], c["model"], c["freqs"], c["P"]
    errs = G2[case + "/noise"]
    assert rel(pplib.get_noise(data, chans=True), errs) < 1e-9
    ps = pplib.fit_phase_shift(data.mean(0), model.mean(0), Ns=100)
    assert abs(ps.phase - G2[case + "/ps.phase"]) < max(0.05 * G2[case + "/ps.phase_err"], 1e-4)   # Nelder-Mead slop
    phi0 = float(G2[case + "/ps.phase"])
    # pplib.fit_portrait (the reference ran TNC)
    r = pplib.fit_portrait(data, model, np.array([phi0, 0.0]), P, freqs, errs=errs)
    g = lambda f: G2[case + "/fp." + f]  # noqa: E731
    assert abs(r.phase - g("phase")) / g("phase_err") < SIG_TOL
    assert abs(r.DM - g("DM")) / g("DM_err") < SIG_TOL
    assert abs(r.chi2 / g("chi2") - 1) < CHI2_TOL
    assert abs(r.red_chi2 / g("red_chi2") - 1) < CHI2_TOL
    assert rel([r.phase_err, r.DM_err], [g("phase_err"), g("DM_err")]) < 1e-4
    assert rel(r.nu_ref, g("nu_ref")) < 1e-4
    assert rel(r.snr, g("snr")) < 1e-6
    assert rel(r.scales, g("scales")) < 1e-4 and rel(r.scale_errs, g("scale_errs")) < 1e-9
    assert r.return_code in (1, 2) and int(g("return_code")) in (0, 1, 2, 4)
    # pptoaslib.fit_portrait_full (trust-ncg)
    r = pptoaslib.fit_portrait_full(data, model, [phi0, 0.0, 0.0, 0.0, 0.0], P, freqs, errs=errs,
                                    fit_flags=[1, 1, 0, 0, 0], log10_tau=False)
    g = lambda f: G2[case + "/full." + f]  # noqa: E731
    assert abs(r.phi - g("phi")) / g("phi_err") < SIG_TOL
    assert abs(r.DM - g("DM")) / g("DM_err") < SIG_TOL
    assert abs(r.chi2 / g("chi2") - 1) < CHI2_TOL
    assert rel([r.phi_err, r.DM_err], [g("phi_err"), g("DM_err")]) < 1e-4
    assert rel(r.nu_DM, g("nu_DM")) < 1e-4
    assert rel(r.scales, g("scales")) < 1e-4 and rel(r.scale_errs, g("scale_errs")) < 1e-4
    assert rel(r.channel_snrs, g("channel_snrs")) < 1e-4
    assert r.return_code == int(g("return_code")) == 2          # trust-ncg's normal exit (pptoaslib.py:1001)


def polished_reference(case, c, flags, log10):
    """The reference's trust-ncg stops when it can no longer predict an improvement of an objective of
    size ~1e7, which leaves some of these weakly constrained fits up to ~2e-3 sigma short of the minimum
    of the reference's own objective (b3_601: 1.8e-3 sigma in alpha).  Exact Newton steps on the oracle's
    restatement of that objective (pinned to the reference's f, gradient and Hessian at 1e-9) from the
    reference's solution, at the reference's output frequencies: the point the reference converges to."""
    from oracle import pp_oracle as orc
    g = lambda f: G2[case + "/full." + f]  # noqa: E731
    nbin = c["data"].shape[1]
    dFT, mFT = orc._spectra(c["data"], c["model"])
    prob = orc._FullProblem(dFT, mFT, G2[case + "/errs"] * np.sqrt(nbin / 2.0), c["P"], c["freqs"], float(g("nu_DM")),
                            float(g("nu_GM")), float(g("nu_tau")), flags, log10)
    x = np.array(g("params"), dtype=np.float64)
    ifit = np.where(flags)[0]
    for _ in range(4):
        x[ifit] -= np.linalg.solve(prob.hess(x)[np.ix_(ifit, ifit)], prob.grad(x)[ifit])
    return x


def run_full(case):
    from pulseportraiture_b200.engine import WidebandPlan
    cfg = G2[case + "/cfg"]
    nchan, nbin, nu0, bw, seed = int(cfg[0]), int(cfg[1]), cfg[2], cfg[3], int(cfg[4])
    tau_s, log10, option, sigma = cfg[5], bool(cfg[6]), int(cfg[7]), cfg[8]
    flags = [int(v) for v in G2[case + "/flags"]]
    c = synth.make_case(nchan, nbin, nu0, bw, seed, tau_data_s=tau_s, sigma=sigma)
    with WidebandPlan(nchan, nbin) as pl:
        pl.set_model(c["model"].astype(np.float32), c["freqs"])
        r = pl.fit_batch(c["data"].astype(np.float32)[None], c["P"], errs=G2[case + "/errs"][None],
                         init=np.array(G2[case + "/init"], dtype=np.float64)[None], fit_flags=flags,
                         log10_tau=log10, option=option)
    g = lambda f: G2[case + "/full." + f]  # noqa: E731
    assert int(r["return_code"][0]) == 0
    xp = polished_reference(case, c, flags, log10)
    # "at the same output reference frequencies" (SURVEY 8c): both sides report tau at their own
    # zero-covariance frequency nu_tau; with alpha ~ -4 a 1e-6 relative difference between the two
    # frequencies moves log10(tau) by 2e-3 sigma, so the device's tau is carried to the reference's nu_tau
    got = r["params"][0].copy()
    if flags[3]:
        ratio = float(g("nu_tau")) / r["nu_out"][0, 2]
        got[3] = got[3] + got[4] * np.log10(ratio) if log10 else got[3] * ratio ** got[4]
    for i, nm in enumerate(["phi", "DM", "GM", "tau", "alpha"]):
        if flags[i]:
            slop = abs(xp[i] - g(nm)) / g(nm + "_err")        # how far the reference stopped from its own minimum
            assert slop < 5e-3, (nm, slop)
            assert abs(got[i] - xp[i]) / g(nm + "_err") < SIG_TOL, nm
            assert abs(got[i] - g(nm)) / g(nm + "_err") < SIG_TOL + slop, nm
            assert rel(r["param_errs"][0, i], g(nm + "_err")) < 1e-4, nm
        else:
            assert abs(r["params"][0, i] - g(nm)) <= 1e-9 * max(1.0, abs(g(nm))), nm
    assert rel(r["nu_out"][0], [g("nu_DM"), g("nu_GM"), g("nu_tau")]) < 1e-4
    assert abs(r["chi2"][0] / g("chi2") - 1) < CHI2_TOL                  # plain 1e-8 relative
    assert rel(r["snr"][0], g("snr")) < 1e-6
    assert rel(r["scales"][0], g("scales")) < 1e-4
    assert rel(r["scale_errs"][0], g("scale_errs")) < 1e-4
    ifit = np.where(flags)[0]
    cm = g("covariance_matrix")
    sc = np.sqrt(np.abs(np.diag(cm)))
    assert np.max(np.abs(r["cov"][0][np.ix_(ifit, ifit)] - cm) / np.outer(sc, sc)) < 1e-3


@pytest.mark.parametrize("case", cases("b3_"))
def test_config3_parity_shapes_against_reference(case):
    run_full(case)


@pytest.mark.parametrize("case", cases("full15_"))
def test_every_flag_pattern_sigma_1p5_against_reference(case):
    run_full(case)


def test_instrumental_response_against_reference():
    from pulseportraiture_b200 import pptoaslib
    nbin = 256
    assert rel(pptoaslib.instrumental_response_FT(nbin, 0.013, 'rect'), G2["ir/rect"]) < 1e-13
    assert np.allclose(np.real(pptoaslib.instrumental_response_FT(nbin, 0.02, 'gauss')), G2["ir/gauss"], rtol=1e-12, atol=1e-300)
    assert np.allclose(pptoaslib.gaussian_profile_FT(nbin, 0.3, 0.05, 2.0), G2["ir/gprof_FT"], rtol=1e-12, atol=1e-300)
    freqs, model = synth.example_model(32, nbin, 1500., 800.)
    model = model.astype(np.float32).astype(np.float64)
    for tag in ("wids", "dm", "both"):
        DM, wids = float(G2["ir_%s/DM" % tag]), list(G2["ir_%s/wids" % tag])
        types = [str(t) for t in G2["ir_%s/types" % tag]]
        resp = pptoaslib.instrumental_response_port_FT(nbin, freqs, DM, synth.P_EXAMPLE, wids, types)
        assert np.allclose(np.real(resp), G2["ir_%s/resp_real" % tag], rtol=1e-12, atol=1e-300)
        conv = pptoaslib.add_instrumental_response(model, freqs, DM, synth.P_EXAMPLE, wids, types)   # device multiply
        ref = G2["ir_%s/conv" % tag]
        assert np.abs(conv - ref).max() < 2e-7 * np.abs(ref).max()                                    # float32 output
    okc = G2["ir_subset/okc"]
    resp = pptoaslib.instrumental_response_port_FT(nbin, freqs[okc], 30.0, synth.P_EXAMPLE, [], [])
    assert np.allclose(np.real(resp), G2["ir_subset/resp_real"], rtol=1e-12, atol=1e-300)
    # the chan_bw override get_TOAs uses for subints with zapped channels gives the same rows
    full = pptoaslib.instrumental_response_port_FT(nbin, freqs, 30.0, synth.P_EXAMPLE, [], [],
                                                   chan_bw=abs(freqs[okc[1]] - freqs[okc[0]]))
    assert np.allclose(np.real(full)[okc], G2["ir_subset/resp_real"], rtol=1e-12, atol=1e-300)


def test_noise_frac_against_reference():
    from pulseportraiture_b200 import pplib
    c = synth.make_case(16, 512, 1500., 800., 701)
    for frac in (1, 2, 4, 8):
        assert rel(pplib.get_noise_PS(c["data"], frac=frac, chans=True), G2["noise/ps_frac%d" % frac]) < 1e-9
    assert rel(pplib.get_noise_PS(c["data"][3], frac=8), G2["noise/ps_prof_frac8"]) < 1e-9
    assert rel(pplib.get_noise(c["data"], method="PS", frac=8, chans=True), G2["noise/ps_frac8"]) < 1e-9


def test_noise_fit_against_reference():
    """get_noise(method='fit') (pplib.py:2255-2284): the find_kc grid search on the device."""
    from pulseportraiture_b200 import pplib
    c = synth.make_case(16, 512, 1500., 800., 701)
    assert rel(pplib.get_noise_fit(c["data"], chans=True), G2["noise/fit_chans"]) < 1e-9
    assert rel(pplib.get_noise(c["data"], method="fit", chans=True), G2["noise/fit_chans"]) < 1e-9
    assert rel(pplib.get_noise_fit(c["data"][3]), G2["noise/fit_prof"]) < 1e-9
    assert rel(pplib.get_noise_fit(c["data"], fact=2.0, chans=True), G2["noise/fit_fact2"]) < 1e-9
    # any nbin
    from oracle import pp_oracle as orc
    c2 = synth.make_case(6, 1000, 1500., 800., 702)
    assert rel(pplib.get_noise_fit(c2["data"], chans=True), orc.get_noise_fit(c2["data"], chans=True)) < 1e-9
