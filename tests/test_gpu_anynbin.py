"""nbin that is not a power of two (the reference transforms rows of any length, pplib.py:2127-2130): the
Bluestein row transforms (csrc/bluestein.cuh) behind every entry point, against outputs of the REFERENCE's
functions (golden_v2.npz, nb_* / nbfull_*) and the oracle."""
import os

import numpy as np
import pytest

from oracle import pp_oracle as orc
from tests import synth
from tests.test_gpu_golden_v2 import G2, cases, rel, run_full, SIG_TOL, CHI2_TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", cases("nb_"))
def test_any_nbin_against_reference(case):
    from pulseportraiture_b200 import pplib, pptoaslib
    nchan, nbin, nu0, bw, seed = G2[case + "/cfg"]
    nchan, nbin, seed = int(nchan), int(nbin), int(seed)
    c = synth.make_case(nchan, nbin, nu0, bw, seed)
    data, model, freqs, P = c["data"], c["model"], c["freqs"], c["P"]
    errs = G2[case + "/noise"]
    assert rel(pplib.get_noise(data, chans=True), errs) < 1e-9
    rot = pplib.rotate_data(data, 0.05, 1e-3, P, freqs, 1400.0)
    assert np.abs(rot[1] - G2[case + "/rot_row1"]).max() < 3e-6 * np.abs(data).max()          # float32 rows
    ps = pplib.fit_phase_shift(data.mean(0), model.mean(0), Ns=100)
    assert ps.lag_index == int(G2[case + "/lag"])                                              # bit-exact lag
    assert abs(ps.phase - G2[case + "/ps.phase"]) < max(0.05 * G2[case + "/ps.phase_err"], 1e-4)
    ex = orc.fit_phase_shift(data.mean(0).astype(np.float32).astype(np.float64),
                             model.mean(0).astype(np.float32).astype(np.float64), Ns=100, polish="exact")
    assert abs(ps.phase - ex.phase) / ex.phase_err < SIG_TOL
    assert rel([ps.scale, ps.snr, ps.red_chi2], [ex.scale, ex.snr, ex.red_chi2]) < 1e-6
    phi0 = float(G2[case + "/ps.phase"])
    r = pplib.fit_portrait(data, model, np.array([phi0, 0.0]), P, freqs, errs=errs)
    g = lambda f: G2[case + "/fp." + f]  # noqa: E731
    assert abs(r.phase - g("phase")) / g("phase_err") < SIG_TOL
    assert abs(r.DM - g("DM")) / g("DM_err") < SIG_TOL
    assert abs(r.chi2 / g("chi2") - 1) < CHI2_TOL and abs(r.red_chi2 / g("red_chi2") - 1) < CHI2_TOL
    assert rel([r.phase_err, r.DM_err, r.nu_ref, r.snr], [g("phase_err"), g("DM_err"), g("nu_ref"), g("snr")]) < 1e-4
    assert rel(r.scales, g("scales")) < 1e-4 and rel(r.scale_errs, g("scale_errs")) < 1e-9
    r = pptoaslib.fit_portrait_full(data, model, [phi0, 0.0, 0.0, 0.0, 0.0], P, freqs, errs=errs,
                                    fit_flags=[1, 1, 0, 0, 0], log10_tau=False)
    g = lambda f: G2[case + "/full." + f]  # noqa: E731
    assert abs(r.phi - g("phi")) / g("phi_err") < SIG_TOL and abs(r.DM - g("DM")) / g("DM_err") < SIG_TOL
    assert abs(r.chi2 / g("chi2") - 1) < CHI2_TOL
    assert rel(r.scale_errs, g("scale_errs")) < 1e-4 and rel(r.channel_snrs, g("channel_snrs")) < 1e-4


@pytest.mark.parametrize("case", cases("nbfull_"))
def test_any_nbin_five_parameters_against_reference(case):
    run_full(case)


@pytest.mark.parametrize("nchan,nbin", [(20, 1000), (12, 1536), (9, 66), (6, 4094)])
def test_any_nbin_batch_with_guess_int16_float64_and_align(nchan, nbin):
    """The whole batch path (noise measured, FFTFIT guess, fit, fused ppalign sum) at other nbin, from
    float32, float64 and int16 inputs, against the oracle."""
    from pulseportraiture_b200 import pptoas
    from pulseportraiture_b200.engine import WidebandPlan
    nsub = 5
    cs = [synth.make_case(nchan, nbin, 1500., 800., 9300 + s) for s in range(nsub)]
    data = np.stack([c["data"] for c in cs])
    P, freqs, model = cs[0]["P"], cs[0]["freqs"], cs[0]["model"]
    with WidebandPlan(nchan, nbin) as pl:
        pl.set_model(model.astype(np.float32), freqs)
        r = pl.fit_batch(data.astype(np.float32), P, align=True, want_chan_sums=False)
        r64 = pl.fit_batch(data, P)
        raw, scl, offs, dec = pptoas.quantize_subints(data)
        r16 = pl.fit_batch(raw, P, dat_scl=scl, dat_offs=offs)
        rdec = pl.fit_batch(dec, P)
        noise = pl.get_noise_batch(data.astype(np.float32))
        # the two-step path: rotate + accumulate with the fitted parameters and weights
        w = r["scales"] / np.where(r["noise"] > 0, r["noise"], np.inf) ** 2
        asum, wsum = pl.align_accumulate(data.astype(np.float32), r["params"][:, 0], r["params"][:, 1], P,
                                         r["nu_out"][:, 0], w)
    for k in r64:
        assert np.array_equal(r64[k], r[k], equal_nan=True) if k not in ("align_sum", "align_wsum") else True, k
        assert np.array_equal(r16[k], rdec[k], equal_nan=True), k
    for s, c in enumerate(cs):
        errs = orc.get_noise(c["data"], chans=True)
        assert rel(noise[s], errs) < 1e-11 and rel(r["noise"][s], errs) < 1e-11
        ref, _, _ = orc.toa_core(c["data"], c["model"], P, freqs, errs, polish="exact")
        assert int(r["lag_index"][s]) == ref.lag_index
        assert abs(r["params"][s, 0] - ref.phi) / ref.phi_err < SIG_TOL
        assert abs(r["params"][s, 1] - ref.DM) / ref.DM_err < SIG_TOL
        assert abs(r["chi2"][s] / ref.chi2 - 1) < CHI2_TOL
        assert rel(r["red_chi2"][s], ref.red_chi2) < CHI2_TOL
        assert rel(r["scales"][s], ref.scales) < 1e-5 and rel(r["scale_errs"][s], ref.scale_errs) < 1e-5
    sc = np.abs(asum).max()
    assert np.abs(r["align_sum"] - asum).max() < 3e-6 * sc and np.allclose(r["align_wsum"], wsum, rtol=1e-12)
    ref_sum = sum(w[s][:, None] * orc.rotate_data(cs[s]["data"], r["params"][s, 0], r["params"][s, 1], P, freqs,
                                                  r["nu_out"][s, 0]) for s in range(nsub))
    assert np.abs(r["align_sum"] - ref_sum).max() < 3e-6 * sc


def test_any_nbin_model_generation_and_gettoas():
    """The facade end to end at nbin = 1000: device model generation, get_TOAs, narrowband TOAs."""
    from pulseportraiture_b200 import pptoas, pplib
    from tests.test_gpu_parity import _fake_archive
    d, cs = _fake_archive(3, 24, 1000, 9400, DM_stored=1e-3)
    gt = pptoas.GetTOAs([d], synth.GMODEL, quiet=True)
    gt.get_TOAs()
    for s, c in enumerate(cs):
        ok = d.ok_ichans[s]
        ref, _, _ = orc.toa_core(c["data"][ok], c["model"][ok], c["P"], c["freqs"][ok], d.noise_stds[s, 0, ok],
                                 weights=d.weights[s, ok], SNRs=d.SNRs[s, 0, ok], DM_stored=1e-3, polish="exact")
        assert abs(gt.phis[0][s] - ref.phi) / ref.phi_err < SIG_TOL
        assert abs(gt.DMs[0][s] / d.doppler_factors[s] - ref.DM) / ref.DM_err < SIG_TOL
    _, _, m_dev = pplib.read_model(synth.GMODEL, pplib.get_bin_centers(1000), cs[0]["freqs"], cs[0]["P"], quiet=True, device=True)
    _, m_ref = synth.example_model(24, 1000, 1500., 800.)
    assert np.abs(m_dev - m_ref).max() < 2e-6 * np.abs(m_ref).max()
    # a scattered model (TAU in the .gmodel -> the scattering multiply of the rotation path)
    freqs2, m_sc = synth.example_model(24, 1000, 1500., 800., tau_s=30e-6)
    gm = pplib.read_model(synth.GMODEL, quiet=True)
    params = np.array(gm[4], dtype=np.float64)
    params[1] = 30e-6 * 1000 / cs[0]["P"]
    m_dev2 = pplib.gen_gaussian_portrait_device(gm[1], params, gm[6], pplib.get_bin_centers(1000), freqs2, gm[2])
    assert np.abs(m_dev2 - m_sc).max() < 3e-6 * np.abs(m_sc).max()


@pytest.mark.parametrize("nchan,nbin", [(16, 1000), (8, 1536), (6, 3000), (10, 120)])
def test_mixed_radix_rows_match_bluestein_rows(nchan, nbin, monkeypatch):
    """nbin/2 = 2^a 3^b 5^c takes the direct mixed-radix transform (bluestein.cuh fft_mixed); PP_FORCE_BLUESTEIN
    keeps such a plan on the chirp-z path: both are the same DFT, so fits, noise and rotated rows agree to
    rounding."""
    from pulseportraiture_b200.engine import WidebandPlan
    nsub = 5
    cs = [synth.make_case(nchan, nbin, 1500., 800., 8800 + s) for s in range(nsub)]
    data = np.stack([c["data"] for c in cs]).astype(np.float32)
    P, freqs, model = cs[0]["P"], cs[0]["freqs"], cs[0]["model"].astype(np.float32)
    out = []
    for force in (False, True):
        if force:
            monkeypatch.setenv("PP_FORCE_BLUESTEIN", "1")
        else:
            monkeypatch.delenv("PP_FORCE_BLUESTEIN", raising=False)
        with WidebandPlan(nchan, nbin) as pl:
            pl.set_model(model, freqs)
            r = pl.fit_batch(data, P)
            rot = pl.rotate_batch(data, np.full(nsub, 0.123), np.full(nsub, 2e-3), P, 1400.0)
            noise = pl.get_noise_batch(data)
            out.append(({k: np.array(v) for k, v in r.items()}, np.array(rot), np.array(noise)))
    (a, ra, na), (b, rb, nb_) = out
    assert (a["return_code"] == 0).all() and (b["return_code"] == 0).all()
    assert np.array_equal(a["lag_index"], b["lag_index"])
    assert np.max(np.abs(a["params"][:, :2] - b["params"][:, :2]) / a["param_errs"][:, :2]) < 1e-6
    assert np.max(np.abs(a["chi2"] / b["chi2"] - 1)) < 1e-11
    assert np.max(np.abs(na / nb_ - 1)) < 1e-12
    assert np.max(np.abs(ra - rb)) < 2e-6 * np.abs(data).max()
    ref = orc.get_noise(cs[0]["data"], chans=True)
    assert rel(na[0], ref) < 1e-9
