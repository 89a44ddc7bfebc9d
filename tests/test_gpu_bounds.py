"""GPU tests of the parameter bounds (the scipy-TNC ``bounds`` of pplib.fit_portrait,
pplib.py:2102-2148, and pptoaslib.fit_portrait_full, pptoaslib.py:1008-1014; defaults of
pptoas.py:461-469): the device Newton solvers treat them as an active set and must land on the
same constrained optimum as the oracle's TNC run."""
import numpy as np
import pytest

from oracle import pp_oracle as orc
from tests import synth

pytestmark = pytest.mark.gpu

SIG_TOL = 1e-3
CHI2_TOL = 1e-8


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def test_fit_portrait_dm_bound_active_and_inactive():
    from pulseportraiture_b200 import pplib
    c = synth.make_case(64, 512, 1500., 800., 4242, phi=0.123, dDM=3e-4)
    free = orc.fit_portrait(c["data"], c["model"], [0.12, 0.0], c["P"], c["freqs"])
    hi = free.DM - 2 * free.DM_err
    for bounds, init in (([(None, None), (None, hi)], [0.12, 0.0]),          # reached from inside
                         ([(None, None), (None, hi)], [0.12, hi + 1e-3]),    # start outside the box
                         ([(0.1231, 0.2), (None, None)], [0.1235, 0.0]),     # phase bound active
                         ([(-0.5, 0.5), (free.DM - 5 * free.DM_err, free.DM + 5 * free.DM_err)], [0.12, 0.0])):
        ref = orc.fit_portrait(c["data"], c["model"], init, c["P"], c["freqs"], bounds=bounds)
        r = pplib.fit_portrait(c["data"], c["model"], init, c["P"], c["freqs"], bounds=bounds)
        assert r.device_return_code == 0 and r.return_code in (1, 2)     # a converged TNC status, pplib.py:2159
        assert abs(r.phase - ref.phase) / ref.phase_err < SIG_TOL
        assert abs(r.DM - ref.DM) / ref.DM_err < SIG_TOL
        assert abs(r.chi2 / ref.chi2 - 1) < CHI2_TOL
        assert rel([r.phase_err, r.DM_err], [ref.phase_err, ref.DM_err]) < 1e-4
        assert rel(r.nu_ref, ref.nu_ref) < 1e-4
        assert rel(r.scales, ref.scales) < 1e-5
        if bounds[1][1] == hi:
            assert abs(r.DM - hi) <= 1e-12 * abs(hi) and ref.DM == hi      # on the bound
            assert r.chi2 > free.chi2 + 3.0                                # 2 sigma away: chi2 + 4
    # the last (inactive) box gives the free optimum
    assert abs(r.phase - free.phase) / free.phase_err < SIG_TOL
    assert abs(r.DM - free.DM) / free.DM_err < SIG_TOL


def test_fit_portrait_full_tnc_bounds():
    from pulseportraiture_b200 import pptoaslib
    c = synth.make_case(64, 512, 600., 400., 4343, phi=0.2, dDM=3e-4, tau_data_s=50e-6, sigma=0.5)
    P = c["P"]
    init = [0.19, 0.0, 0.0, np.log10(0.8 * 50e-6 / P), -4.0]
    kw = dict(fit_flags=[1, 1, 0, 1, 1], log10_tau=True, method="TNC")
    free = orc.fit_portrait_full(c["data"], c["model"], init, P, c["freqs"], **kw)
    lo_alpha = free.alpha + 2 * free.alpha_err
    # tau is referenced to nu_fit inside the fit: a bound 1.5 sigma below the free optimum *there*
    nu_fit = c["freqs"].mean()
    tau_fit_free = free.tau + free.alpha * np.log10(nu_fit / free.nu_tau)
    boxes = [
        [(None, None), (None, None), (None, None), (None, None), (lo_alpha, 10.0)],
        [(None, None), (None, None), (None, None), (None, tau_fit_free - 2e-3), (-10.0, 10.0)],
        [(None, None), (None, free.DM - 2 * free.DM_err), (None, None), (None, None), (lo_alpha, 10.0)],
        [(None, None), (None, None), (None, None), (np.log10((10 * 512) ** -1.0), None), (-10.0, 10.0)],  # pptoas.py:461-469
    ]
    for ib, bounds in enumerate(boxes):
        ref = orc.fit_portrait_full(c["data"], c["model"], init, P, c["freqs"], bounds=bounds, **kw)
        r = pptoaslib.fit_portrait_full(c["data"], c["model"], init, P, c["freqs"], bounds=bounds, **kw)
        assert r.device_return_code == 0 and r.return_code in (0, 1, 2), ib   # scipy's converged statuses
        for i, nm in ((0, "phi"), (1, "DM"), (3, "tau"), (4, "alpha")):
            assert abs(r[nm] - ref[nm]) / ref[nm + "_err"] < SIG_TOL, (ib, nm)
            assert rel(r[nm + "_err"], ref[nm + "_err"]) < 1e-4, (ib, nm)
        assert abs(r.chi2 - ref.chi2) <= CHI2_TOL * max(ref.chi2, ref.snr ** 2), ib
        assert rel([r.nu_DM, r.nu_tau], [ref.nu_DM, ref.nu_tau]) < 1e-4, ib
        if ib in (0, 2):
            assert ref.alpha == lo_alpha and abs(r.alpha - lo_alpha) < 1e-12
        if ib == 2:
            assert abs(r.DM - bounds[1][1]) <= 1e-12 * abs(r.DM)
        if ib == 3:                                       # the default TNC box is not active here
            for i, nm in ((0, "phi"), (1, "DM"), (3, "tau"), (4, "alpha")):
                assert abs(r[nm] - free[nm]) / free[nm + "_err"] < SIG_TOL
    # other methods ignore bounds, as the reference does (bounds reach scipy for TNC only)
    r = pptoaslib.fit_portrait_full(c["data"], c["model"], init, P, c["freqs"], bounds=boxes[0],
                                    fit_flags=[1, 1, 0, 1, 1], log10_tau=True, method="trust-ncg")
    assert abs(r.alpha - free.alpha) / free.alpha_err < SIG_TOL


def test_batch_with_guess_and_bounds():
    """pp_fit_batch with the FFTFIT guess and a DM box that is active for some subints only."""
    from pulseportraiture_b200 import engine
    nsub, nchan, nbin, nu0, bw = 8, 32, 1024, 1500., 800.
    cases = [synth.make_case(nchan, nbin, nu0, bw, 9300 + s) for s in range(nsub)]
    data = np.stack([c["data"] for c in cases]).astype(np.float32)
    P, freqs = cases[0]["P"], cases[0]["freqs"]
    errs = np.stack([orc.get_noise(c["data"], chans=True) for c in cases])
    with engine.WidebandPlan(nchan, nbin) as pl:
        pl.set_model(cases[0]["model"].astype(np.float32), freqs)
        free = pl.fit_batch(data, P, errs=errs)
        hi = float(np.median(free["params"][:, 1]))
        bounds = [(None, None), (None, hi)]
        r = pl.fit_batch(data, P, errs=errs, bounds=bounds)
    nact = 0
    for s, c in enumerate(cases):
        init = [r["phi_guess"][s], 0.0, 0.0, 0.0, 0.0]
        ref = orc.fit_portrait_full(c["data"], c["model"], init, P, freqs, errs=errs[s], nu_fits=[freqs.mean()] * 3,
                                    fit_flags=[1, 1, 0, 0, 0], log10_tau=False, method="TNC",
                                    bounds=bounds + [(None, None)] * 3)
        assert int(r["return_code"][s]) == 0
        assert abs(r["params"][s, 0] - ref.phi) / ref.phi_err < SIG_TOL
        assert abs(r["params"][s, 1] - ref.DM) / ref.DM_err < SIG_TOL
        assert abs(r["chi2"][s] / ref.chi2 - 1) < CHI2_TOL
        assert r["params"][s, 1] <= hi * (1 + 1e-12)
        if free["params"][s, 1] > hi:
            nact += 1
            assert abs(r["params"][s, 1] - hi) <= 1e-12 * abs(hi)
        else:
            assert abs(r["params"][s, 1] - free["params"][s, 1]) / ref.DM_err < SIG_TOL
    assert 2 <= nact <= 6


def test_gettoas_tnc_default_bounds():
    """get_TOAs(method='TNC') installs the reference's default box (pptoas.py:461-469); with a
    normal scattering fit it is not active and the TOAs equal the default method's."""
    from pulseportraiture_b200 import pptoas
    from tests.test_gpu_parity import _fake_archive
    tau_s = 50e-6
    data, cases = _fake_archive(3, 64, 512, 8100, tau_s=tau_s, nu0=600., bw=400., sigma=0.5)
    out = []
    for method in ("trust-ncg", "TNC"):
        gt = pptoas.GetTOAs([data], synth.GMODEL, quiet=True)
        gt.get_TOAs(fit_scat=True, log10_tau=True, scat_guess=(0.8 * tau_s, 600., -4.0), bary=False, method=method)
        out.append(gt)
    for s in range(3):
        assert abs(out[0].phis[0][s] - out[1].phis[0][s]) < SIG_TOL * out[0].phi_errs[0][s]
        assert abs(out[0].taus[0][s] - out[1].taus[0][s]) < SIG_TOL * out[0].tau_errs[0][s]
    # a tight alpha box through the facade is honoured
    gt = pptoas.GetTOAs([data], synth.GMODEL, quiet=True)
    gt.get_TOAs(fit_scat=True, log10_tau=True, scat_guess=(0.8 * tau_s, 600., -4.0), bary=False, method="TNC",
                bounds=[(None, None), (None, None), (None, None), (None, None), (-3.5, 10.0)])
    assert np.all(np.abs(gt.alphas[0] + 3.5) < 1e-12)
