"""Round-2 additions of the drop-in facade, on the GPU: dedispersed archives (dmc), tscrunch, the
instrumental response in get_TOAs, nu_fits, scipy return codes, float64 input without a host pass,
ppalign on another frequency grid with several data channels per template channel and on Stokes data,
.npz round trips, print_paz_cmds."""
import numpy as np
import pytest

from oracle import pp_oracle as orc
from tests import synth
from tests.test_gpu_parity import _fake_archive, rel, SIG_TOL, CHI2_TOL

pytestmark = pytest.mark.gpu


def test_float64_input_needs_no_host_pass_and_is_bit_identical():
    """PP_DATA_F64: float64 portraits (the reference's array type) are rounded to float32 on the device;
    host numpy arrays and device tensors give exactly what the float32 arrays give."""
    import torch
    from pulseportraiture_b200.engine import WidebandPlan
    cases = [synth.make_case(32, 512, 1500., 800., 4400 + s) for s in range(70)]
    d64 = np.stack([c["data"] for c in cases]) + 1e-9          # not float32-representable any more
    d32 = d64.astype(np.float32)
    with WidebandPlan(32, 512) as pl:
        pl.set_model(cases[0]["model"].astype(np.float32), cases[0]["freqs"])
        pl.set_chunk(16)                                       # several staged chunks
        r32 = pl.fit_batch(d32, cases[0]["P"])
        r64 = pl.fit_batch(d64, cases[0]["P"])
        r64d = pl.fit_batch(torch.from_numpy(d64).cuda(), cases[0]["P"])
    for k in r32:
        assert np.array_equal(r32[k], r64[k], equal_nan=True), k
        assert np.array_equal(r32[k], r64d[k], equal_nan=True), k
    # chunks of >= 8 MB from pageable memory: the host threads that stage them round to float32 on the way
    c = synth.make_case(64, 2048, 1500., 800., 4500)
    rng = np.random.RandomState(3)
    d64 = c["data"][None] + rng.normal(0.0, 1.0, (40,) + c["data"].shape) + 1e-9
    d32 = d64.astype(np.float32)
    with WidebandPlan(64, 2048) as pl:
        pl.set_model(c["model"].astype(np.float32), c["freqs"])
        pl.set_chunk(16)
        r32 = pl.fit_batch(d32, c["P"])
        r64 = pl.fit_batch(d64, c["P"])
    for k in r32:
        assert np.array_equal(r32[k], r64[k], equal_nan=True), k


def test_gettoas_dedispersed_archive_is_redispersed():
    """dmc = 1: get_TOAs restores the dispersion of the stored DM before fitting, as the reference's
    reload with dededisperse=True does (pptoas.py:256-265); same TOAs as from the dispersed archive."""
    from pulseportraiture_b200 import pptoas, pplib
    DM = 2.5e-3
    d, cases = _fake_archive(4, 32, 512, 8300, DM_stored=DM)
    g0 = pptoas.GetTOAs([d], synth.GMODEL, quiet=True)
    g0.get_TOAs()
    dd = pptoas.DataBunch(**dict(d))
    # what PSRCHIVE stores after arch.dedisperse(): delays removed about the centre frequency
    dd["subints"] = pplib.rotate_data(np.asarray(d.subints), 0.0, DM, np.asarray(d.Ps), np.asarray(d.freqs), d.nu0)
    dd["dmc"] = 1
    g1 = pptoas.GetTOAs([dd], synth.GMODEL, quiet=True)
    g1.get_TOAs()
    for s in range(4):
        # two float32 rotations apart: far inside the parameter errors
        assert abs(g1.phis[0][s] - g0.phis[0][s]) / g0.phi_errs[0][s] < 2e-2
        assert abs(g1.DMs[0][s] - g0.DMs[0][s]) / g0.DM_errs[0][s] < 2e-2
        assert abs(g1.DMs[0][s] - DM) < 5e-3                  # not the ~0 a fit of dedispersed data would give
    # and against the oracle on the re-dispersed portraits themselves
    back = pptoas.restore_dispersion(dd)
    s = 2
    ok = d.ok_ichans[s]
    port = np.asarray(back.subints)[s, 0].astype(np.float32).astype(np.float64)
    ref, _, _ = orc.toa_core(port[ok], cases[s]["model"][ok], cases[s]["P"], cases[s]["freqs"][ok],
                             d.noise_stds[s, 0, ok], weights=d.weights[s, ok], SNRs=d.SNRs[s, 0, ok],
                             DM_stored=DM, polish="exact")
    assert abs(g1.phis[0][s] - ref.phi) / ref.phi_err < SIG_TOL
    assert abs(g1.DMs[0][s] / d.doppler_factors[s] - ref.DM) / ref.DM_err < SIG_TOL


def test_gettoas_tscrunch():
    """tscrunch=True: one TOA per archive from the weighted average of its (already aligned) subints."""
    from pulseportraiture_b200 import pptoas
    nsub = 6
    c0 = synth.make_case(32, 512, 1500., 800., 8400, phi=0.21, dDM=4e-4, sigma=0.0)
    rng = np.random.RandomState(3)
    clean = c0["data"]                                          # noiseless; independent noise per subint
    subs = np.stack([clean + rng.normal(0, 1.5, clean.shape) for _ in range(nsub)])
    d, _ = _fake_archive(nsub, 32, 512, 8400)
    d["subints"] = subs[:, None]
    d["noise_stds"] = np.stack([orc.get_noise(x, chans=True) for x in subs])[:, None]
    d["ok_ichans"] = [np.arange(32)] * nsub
    d["weights"] = np.tile(np.linspace(0.5, 1.5, 32), (nsub, 1))
    gt = pptoas.GetTOAs([d], synth.GMODEL, quiet=True)
    gt.get_TOAs(tscrunch=True, bary=False)
    assert len(gt.TOA_list) == 1 and len(gt.phis[0]) == 1
    ts = pptoas.tscrunch_databunch(d)
    avg = np.average(subs, axis=0, weights=np.ones(nsub))       # equal weights per channel across subints
    assert np.allclose(np.asarray(ts.subints)[0, 0], avg, atol=1e-12)
    port = avg.astype(np.float32).astype(np.float64)
    errs = np.asarray(ts.noise_stds)[0, 0]
    assert rel(errs, orc.get_noise(port, chans=True)) < 1e-6
    ref, _, _ = orc.toa_core(port, c0["model"], c0["P"], c0["freqs"], errs, weights=ts.weights[0],
                             SNRs=ts.SNRs[0, 0], polish="exact")
    assert abs(gt.phis[0][0] - ref.phi) / ref.phi_err < SIG_TOL
    assert abs(gt.DMs[0][0] - ref.DM) / ref.DM_err < SIG_TOL
    g_all = pptoas.GetTOAs([d], synth.GMODEL, quiet=True)
    g_all.get_TOAs(bary=False)
    assert gt.phi_errs[0][0] < 0.6 * np.median(g_all.phi_errs[0])      # sqrt(6) more signal to noise


def test_gettoas_instrumental_response_and_nu_fits():
    """add_instrumental_response=True: the model is multiplied by instrumental_response_port_FT on the
    usable channels' frequencies with the subint's own period (pptoas.py:388-394), one model per period;
    nu_fits is filled (pptoas.py:407); return codes are scipy's."""
    from pulseportraiture_b200 import pptoas
    d, cases = _fake_archive(3, 32, 256, 8500)
    d["Ps"] = np.array([c["P"] for c in cases]) * np.array([1.0, 1.0 + 3e-6, 1.0 - 2e-6])
    gt = pptoas.GetTOAs([d], synth.GMODEL, quiet=True)
    gt.ird = gt.instrumental_response_dict = {'DM': 30.0, 'wids': [0.004], 'irf_types': ['gauss']}
    gt.get_TOAs(add_instrumental_response=True, bary=False)
    for s, c in enumerate(cases):
        ok = d.ok_ichans[s]
        P = d.Ps[s]
        freqs = c["freqs"]
        _, model = synth.example_model(32, 256, 1500., 800.)
        resp = orc.instrumental_response_port_FT(256, freqs[ok], 30.0, P, [0.004], ['gauss'])
        modelx = np.fft.irfft(resp * np.fft.rfft(model[ok], axis=-1), axis=-1)
        ref, _, _ = orc.toa_core(c["data"][ok], modelx.astype(np.float32).astype(np.float64), P, freqs[ok],
                                 d.noise_stds[s, 0, ok], weights=d.weights[s, ok], SNRs=d.SNRs[s, 0, ok],
                                 polish="exact")
        assert abs(gt.phis[0][s] - ref.phi) / ref.phi_err < 5e-3      # the model passes through float32 twice
        assert abs(gt.DMs[0][s] - ref.DM) / ref.DM_err < 5e-3
        assert rel(gt.nu_fits[0][s], [orc.guess_fit_freq(freqs[ok], d.SNRs[s, 0, ok])] * 3) < 1e-14
        assert gt.rcs[0][s] == 2                                      # trust-ncg's normal exit
    g2 = pptoas.GetTOAs([d], synth.GMODEL, quiet=True)
    g2.get_TOAs(method='TNC', bary=False)
    assert np.all(g2.rcs[0] == 1)                                     # TNC: FCONVERGED
    assert np.any(np.abs(g2.phis[0] - gt.phis[0]) > 0)                # the response changes the template


def test_scipy_return_codes():
    from pulseportraiture_b200 import pplib, pptoaslib
    assert pplib.scipy_return_code(0, 'TNC') == 1 and pplib.scipy_return_code(1, 'TNC') == 3
    assert pplib.scipy_return_code(3, 'TNC') not in (1, 2, 4)          # a non-finite fit is not benign
    assert list(pplib.scipy_return_code(np.array([0, 1, 3]), 'trust-ncg')) == [2, 1, 3]
    c = synth.make_case(16, 256, 1500., 800., 8600)
    bad = c["data"].copy()
    bad[:, 7] = np.nan
    r = pptoaslib.fit_portrait_full(bad, c["model"], [0.0, 0.0, 0.0, 0.0, 0.0], c["P"], c["freqs"],
                                    fit_flags=[1, 1, 0, 0, 0], log10_tau=False)
    assert r.return_code == pplib.scipy_return_code(r.device_return_code, 'trust-ncg')
    assert np.all(r.scales == 0)                                        # no usable channel: nothing was fit
    r = pptoaslib.fit_portrait_full(c["data"], c["model"], [c["phi"], 0.0, 0.0, 0.0, 0.0], c["P"], c["freqs"],
                                    fit_flags=[1, 1, 0, 0, 0], log10_tau=False, method='Newton-CG')
    assert r.device_return_code == 0 and r.return_code == 0


def _align_oracle(archives, template, tfreqs, niter=1, last_wins=True):
    """Restatement of ppalign.py:113-213 for test use (same-grid or nearest-channel mapping, total
    intensity alignment, every polarisation accumulated with numpy's fancy-index +=)."""
    model = template.copy()
    npol = np.asarray(archives[0].subints).shape[1]
    nchan, nbin = template.shape
    for _ in range(niter):
        aligned = np.zeros((npol, nchan, nbin))
        tw = np.zeros((nchan, nbin))
        for d in archives:
            for isub in d.ok_isubs:
                ichans = np.asarray(d.ok_ichans[isub])
                mi = np.array([np.argmin(abs(tfreqs - d.freqs[isub, c])) for c in ichans])
                port = np.asarray(d.subints)[isub, 0, ichans]
                freqs = d.freqs[isub, ichans]
                errs = d.noise_stds[isub, 0, ichans]
                P = d.Ps[isub]
                nu_fit = orc.guess_fit_freq(freqs, d.SNRs[isub, 0, ichans])
                mod = model[mi].astype(np.float32).astype(np.float64)
                g = orc.fit_phase_shift(np.average(port, axis=0, weights=d.weights[isub, ichans]), mod.mean(axis=0),
                                        Ns=nbin, polish="exact")
                r = orc.fit_portrait_full(port, mod, [g.phase, 0.0, 0.0, 0.0, 0.0], P, freqs, [nu_fit] * 3,
                                          [None] * 3, errs, [1, 1, 0, 0, 0], log10_tau=False)
                w = np.outer(r.scales / errs ** 2, np.ones(nbin))
                for ipol in range(npol):
                    aligned[ipol, mi] += w * orc.rotate_data(np.asarray(d.subints)[isub, ipol, ichans], r.phi, r.DM,
                                                             P, freqs, r.nu_DM)
                tw[mi] += w
        good = tw[:, 0] > 0
        aligned[:, good] /= tw[good][None]
        model = aligned[0]
    return aligned, tw[:, 0]


def test_ppalign_many_data_channels_per_template_channel_and_stokes():
    """An archive with twice the template's channels (two data channels map onto each template channel:
    numpy's fancy-index += keeps the last one per subint, ppalign.py:205-209) and four polarisations
    aligned with the total-intensity fit."""
    from pulseportraiture_b200 import ppalign, pptoas
    nsub, nbin = 3, 256
    tfreqs, template = synth.example_model(16, nbin, 1500., 800.)
    d, cases = _fake_archive(nsub, 32, nbin, 8700)
    rng = np.random.RandomState(9)
    I = np.asarray(d.subints)[:, 0]
    stokes = np.stack([I, 0.3 * I + rng.normal(0, 1, I.shape), -0.2 * I + rng.normal(0, 1, I.shape),
                       0.1 * I + rng.normal(0, 1, I.shape)], axis=1)
    d["subints"] = stokes.astype(np.float32).astype(np.float64)
    d["npol"], d["state"] = 4, "Stokes"
    tmpl = pptoas.DataBunch(**dict(d))
    tmpl.update(subints=template[None, None], freqs=tfreqs[None], ok_ichans=[np.arange(16)], ok_isubs=np.array([0]),
                nchan=16, nsub=1, masks=None)
    out = ppalign.align_archives([d], tmpl, fit_dm=True, pscrunch=False, niter=1, quiet=True)
    assert out.port.shape == (4, 16, nbin)
    ref, tw = _align_oracle([d], template, tfreqs)
    scale = np.abs(ref[0]).max()
    for ipol in range(4):
        assert np.abs(out.port[ipol] - ref[ipol]).max() < 2e-5 * scale, ipol
    assert rel(out.weights[tw > 0], tw[tw > 0]) < 1e-4
    # total intensity only (pscrunch): the same first plane
    out1 = ppalign.align_archives([d], tmpl, fit_dm=True, pscrunch=True, niter=1, quiet=True)
    assert out1.port.shape == (16, nbin)
    assert np.abs(out1.port - ref[0]).max() < 2e-5 * scale


def test_align_archives_from_npz_paths(tmp_path):
    """archives given as .npz paths (no prof_SNR stored) must load and align."""
    from pulseportraiture_b200 import ppalign, pptoas
    d, cases = _fake_archive(3, 16, 256, 8800)
    d["prof_SNR"] = None
    path = str(tmp_path / "a.npz")
    pptoas.save_databunch(path, d)
    back = pptoas.load_data(path)
    assert back.prof_SNR is None and back.prof is None and back.flux_prof is None
    a = ppalign.align_archives([path], cases[0]["model"], niter=1, quiet=True)
    b = ppalign.align_archives([d], cases[0]["model"], niter=1, quiet=True)
    assert np.array_equal(a.port, b.port)
    d["prof_SNR"] = 3.0
    pptoas.save_databunch(path, d)
    assert pptoas.load_data(path).prof_SNR == 3.0
    c = ppalign.align_archives([path], cases[0]["model"], niter=1, SNR_cutoff=10.0, quiet=True)
    assert not np.any(c.weights)                                # cut: nothing averaged


def test_print_paz_cmds_mirrors_reference(tmp_path):
    from pulseportraiture_b200 import ppzap
    zl = [[[3, 5], []], [[], [7]]]
    lines = ppzap.print_paz_cmds(["a.fits", "b.ar"], zl, all_subs=False, modify=False, quiet=True,
                                 outfile=str(tmp_path / "z.sh"))
    assert lines == ["paz -e zap a.fits", "paz -m -I -z 3 -w 0 a.zap", "paz -m -I -z 5 -w 0 a.zap",
                     "paz -e zap b.ar", "paz -m -I -z 7 -w 1 b.zap"]
    assert open(str(tmp_path / "z.sh")).read().splitlines() == lines
    lines = ppzap.print_paz_cmds(["a.fits"], [[[3], [3], [4]]], all_subs=True, modify=True, quiet=True,
                                 outfile=str(tmp_path / "y.sh"))
    assert lines == ["paz -m -z 3 a.fits", "paz -m -z 4 a.fits"]
    assert ppzap.print_paz_cmds([], [], quiet=True) is None


def test_general_solver_coarse_stage_same_optimum():
    """The general solver's coarse-to-fine start (low harmonics of a channel subset first: ppb200.h
    pp_plan_set_coarse) changes the path, not the answer: parameters within 1e-3 sigma and chi2 / errors to
    rounding of the plain Newton iterations, with fewer full-resolution passes; masked channels included."""
    from pulseportraiture_b200 import engine
    nsub, nchan, nbin, nu0, bw, tau_s = 6, 256, 1024, 600., 400., 50e-6
    cases = [synth.make_case(nchan, nbin, nu0, bw, 9100 + s, tau_data_s=tau_s) for s in range(nsub)]
    data = np.stack([c["data"] for c in cases]).astype(np.float32)
    P, freqs = cases[0]["P"], cases[0]["freqs"]
    mask = np.ones((nsub, nchan), dtype=np.uint8)
    mask[1, ::4] = 0            # the first level's channel subset (every 2nd channel here) loses half of its channels
    mask[2, 10:90] = 0
    scat = np.tile([0.8 * tau_s / P * (freqs.mean() / nu0) ** -4.0, -4.0], (nsub, 1))
    for flags in ((1, 1, 0, 1, 1), (1, 1, 1, 1, 1), (1, 1, 0, 1, 0)):
        out = {}
        for frac in (0.0, 0.99):
            with engine.WidebandPlan(nchan, nbin) as pl:
                pl.set_model(cases[0]["model"].astype(np.float32), freqs)
                pl.set_coarse(frac)
                pl.enable_timing(True)
                r = pl.fit_batch(data, P, chan_mask=mask, fit_flags=flags, log10_tau=True, scat_guess=scat)
                out[frac] = ({k: np.array(v) for k, v in r.items() if isinstance(v, np.ndarray)}, pl.stats())
        (a, sa), (b, sb) = out[0.0], out[0.99]
        assert sa["coarse_launches"] == 0 and sb["coarse_launches"] > 0
        assert sb["pass_launches"] < sa["pass_launches"]
        assert (a["return_code"] == 0).all() and (b["return_code"] == 0).all()
        fit = np.array(flags, bool)
        assert np.max(np.abs(a["params"] - b["params"])[:, fit] / a["param_errs"][:, fit]) < 1e-3
        assert np.max(np.abs(b["param_errs"][:, fit] / a["param_errs"][:, fit] - 1)) < 1e-5
        assert np.max(np.abs(b["chi2"] / a["chi2"] - 1)) < 1e-9
        ok = mask.astype(bool)
        assert np.max(np.abs(b["scales"][ok] / a["scales"][ok] - 1)) < 1e-4
    # set_model_steps(1) (every step evaluated on the data) turns the coarse stage off as well
    with engine.WidebandPlan(nchan, nbin) as pl:
        pl.set_model(cases[0]["model"].astype(np.float32), freqs)
        pl.set_model_steps(1)
        pl.fit_batch(data, P, fit_flags=(1, 1, 0, 1, 1), log10_tau=True, scat_guess=scat)
        assert pl.stats()["coarse_launches"] == 0
    with pytest.raises(Exception):
        with engine.WidebandPlan(nchan, nbin) as pl:
            pl.set_coarse(1.5)


@pytest.mark.parametrize("nchan,nbin,log10_tau,tau_s", [(32, 256, True, 400e-6), (48, 512, False, 200e-6),
                                                        (160, 2048, True, 100e-6), (64, 1024, False, 0.0)])
def test_general_solver_coarse_stage_small_shapes(nchan, nbin, log10_tau, tau_s):
    """The coarse levels at the corners of their rules: nbin = 256 (8 harmonic groups: at most 4 may be coarse),
    fewer than 128 channels (no channel stride), linear tau, and a fit that starts at tau = 0 (scattering
    derivatives off, pptoaslib.py:325-330): same optimum as the plain iterations in every case."""
    from pulseportraiture_b200 import engine
    nsub, nu0, bw = 4, 600., 200.
    cases = [synth.make_case(nchan, nbin, nu0, bw, 9300 + s, tau_data_s=tau_s, sigma=0.5) for s in range(nsub)]
    data = np.stack([c["data"] for c in cases]).astype(np.float32)
    P, freqs = cases[0]["P"], cases[0]["freqs"]
    tau0 = 0.8 * tau_s / P * (freqs.mean() / nu0) ** -4.0
    flags = (1, 1, 0, 1, 1) if tau_s else (1, 1, 1, 0, 0)
    scat = np.tile([tau0, -4.0], (nsub, 1))
    out = {}
    for frac in (0.0, 0.99):
        with engine.WidebandPlan(nchan, nbin) as pl:
            pl.set_model(cases[0]["model"].astype(np.float32), freqs)
            pl.set_coarse(frac)
            r = pl.fit_batch(data, P, fit_flags=flags, log10_tau=log10_tau and tau_s > 0, scat_guess=scat)
            out[frac] = ({k: np.array(v) for k, v in r.items() if isinstance(v, np.ndarray)}, pl.stats())
    (a, sa), (b, sb) = out[0.0], out[0.99]
    assert sa["coarse_launches"] == 0
    assert (a["return_code"] == 0).all() and (b["return_code"] == 0).all()
    fit = np.array(flags, bool)
    assert np.max(np.abs(a["params"] - b["params"])[:, fit] / a["param_errs"][:, fit]) < 1e-3
    assert np.max(np.abs(b["param_errs"][:, fit] / a["param_errs"][:, fit] - 1)) < 1e-5
    assert np.max(np.abs(b["chi2"] / a["chi2"] - 1)) < 1e-9


def test_model_harmonic_cutoff_changes_nothing_measurable():
    """The harmonic cut-off (ppb200.h pp_plan_set_model_cutoff): harmonics where a float64 analytic model has no
    power are neither stored nor streamed.  Against the same plan with the cut-off off: chi2 within 1e-9, parameters
    within 1e-6 sigma, errors / scales to rounding, for the (phi, DM) and the general solver; a template with a
    noise floor keeps every harmonic; a float32 copy of the model (rounding floor) keeps far more than the
    float64 one."""
    from pulseportraiture_b200 import engine
    nsub, nchan, nbin, nu0, bw, tau_s = 6, 64, 2048, 1500., 800., 20e-6
    cases = [synth.make_case(nchan, nbin, nu0, bw, 9500 + s, tau_data_s=tau_s) for s in range(nsub)]
    data = np.stack([c["data"] for c in cases]).astype(np.float32)
    P, freqs = cases[0]["P"], cases[0]["freqs"]
    _, model = synth.example_model(nchan, nbin, nu0, bw)        # float64 as generated (make_case rounds its copy)
    scat = np.tile([0.8 * tau_s / P * (freqs.mean() / nu0) ** -4.0, -4.0], (nsub, 1))
    mask = np.ones((nsub, nchan), dtype=np.uint8)
    mask[1, 5:20] = 0
    for kw in (dict(), dict(fit_flags=(1, 1, 0, 1, 1), log10_tau=True, scat_guess=scat)):
        out = {}
        for eps in (0.0, 1e-10):
            with engine.WidebandPlan(nchan, nbin) as pl:
                pl.set_model_cutoff(eps)
                pl.set_model(model, freqs)                      # float64
                r = pl.fit_batch(data, P, chan_mask=mask, **kw)
                out[eps] = ({k: np.array(v) for k, v in r.items()}, pl.stats()["x_keep_frac"])
        (a, ka), (b, kb) = out[0.0], out[1e-10]
        assert ka == 1.0 and 0.05 < kb < 0.6
        assert (a["return_code"] == 0).all() and (b["return_code"] == 0).all()
        fit = a["param_errs"][0] > 0
        assert np.max(np.abs(a["params"] - b["params"])[:, fit] / a["param_errs"][:, fit]) < 1e-6
        assert np.max(np.abs(b["chi2"] / a["chi2"] - 1)) < 1e-9
        assert np.max(np.abs(b["param_errs"][:, fit] / a["param_errs"][:, fit] - 1)) < 1e-8
        ok = mask.astype(bool)
        assert np.max(np.abs(b["scales"][ok] / a["scales"][ok] - 1)) < 1e-7
        assert np.array_equal(a["lag_index"], b["lag_index"])
    with engine.WidebandPlan(nchan, nbin) as pl:
        pl.set_model(model.astype(np.float32), freqs)           # float32: rounding floor in every harmonic
        k32 = pl.stats()["x_keep_frac"]
        pl.set_model(model + np.random.RandomState(1).normal(0.0, 1e-3, model.shape), freqs)   # a noisy template
        knoisy = pl.stats()["x_keep_frac"]
        pl.set_model(model, freqs)
        k64 = pl.stats()["x_keep_frac"]
        pl.set_model_cutoff(0.0)                                # re-evaluated for the model already set
        assert pl.stats()["x_keep_frac"] == 1.0
    assert knoisy == 1.0 and k64 < 0.6 and k32 > k64
    with pytest.raises(Exception):
        with engine.WidebandPlan(nchan, nbin) as pl:
            pl.set_model_cutoff(0.5)
