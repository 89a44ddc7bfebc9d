"""GPU tests of the 'next' rows of SURVEY 8(f): the ppalign inner loop (rotate-accumulate,
iterated template) and the zap scans, against compositions of oracle functions."""
import numpy as np
import pytest

from oracle import pp_oracle as orc
from tests import synth
from tests.test_gpu_parity import _fake_archive, rel, SIG_TOL

pytestmark = pytest.mark.gpu


def oracle_rotate_full(port, phi, DM, GM, freqs, nu_DM, nu_GM, P):
    """pptoaslib.rotate_portrait_full (pptoaslib.py:52-81) restated with numpy."""
    FT = np.fft.rfft(port, axis=-1)
    k = np.arange(FT.shape[-1])
    th = phi + orc.Dconst * DM * (freqs ** -2 - nu_DM ** -2) / P + \
        orc.Dconst ** 2 * GM * (freqs ** -4 - nu_GM ** -4) / P
    return np.fft.irfft(FT * np.exp(2.0j * np.pi * np.outer(th, k)), axis=-1)


def test_rotate_portrait_full_and_scales_full():
    from pulseportraiture_b200 import pptoaslib
    c = synth.make_case(32, 512, 600., 400., 9001, tau_data_s=50e-6, sigma=0.5)
    data, model, freqs, P = c["data"], c["model"], c["freqs"], c["P"]
    out = pptoaslib.rotate_portrait_full(data, 0.21, 1.5e-3, 2.0e-7, freqs, 650., 620., P)
    ref = oracle_rotate_full(data, 0.21, 1.5e-3, 2.0e-7, freqs, 650., 620., P)
    assert np.max(np.abs(out - ref)) < 2e-5 * np.max(np.abs(ref))
    params = [0.1, 2e-4, 1e-8, np.log10(3e-3), -4.2]
    errs = orc.get_noise(data, chans=True)
    sc = pptoaslib.get_scales_full(params, data, model, P, freqs, 610., 590., 600., True, errs=errs)
    dFT, mFT = orc._spectra(data, model)
    prob = orc._FullProblem(dFT, mFT, errs * np.sqrt(256.), P, freqs, 610., 590., 600.,
                            [1, 1, 1, 1, 1], True)
    pr = prob.primitives(params, order=0)
    assert rel(sc, pr["C"] / pr["S"]) < 1e-6


def test_align_accumulate_vs_oracle():
    from pulseportraiture_b200.engine import WidebandPlan
    nsub, nchan, nbin = 7, 48, 1024
    rng = np.random.RandomState(3)
    cases = [synth.make_case(nchan, nbin, 1500., 800., 9100 + s) for s in range(nsub)]
    data = np.stack([c["data"] for c in cases]).astype(np.float32)
    P, freqs = cases[0]["P"], cases[0]["freqs"]
    phi, DM = rng.uniform(-0.5, 0.5, nsub), rng.normal(0, 1e-3, nsub)
    nu_ref = rng.uniform(1200., 1800., nsub)
    w = rng.uniform(0.5, 2.0, (nsub, nchan))
    w[2, 5] = 0.0
    w[4, :] = 0.0
    w[1, 7] = -0.3          # a negative fitted amplitude weighs in with its sign (ppalign.py:202-209)
    with WidebandPlan(nchan, nbin) as pl:
        pl.set_freqs(freqs)
        acc, wsum = pl.align_accumulate(data, phi, DM, P, nu_ref, w)
    ref = np.zeros((nchan, nbin))
    for s in range(nsub):
        ref += w[s][:, None] * orc.rotate_data(data[s].astype(np.float64), phi[s], DM[s], P, freqs, nu_ref[s])
    assert rel(wsum, w.sum(axis=0)) < 1e-14
    assert np.max(np.abs(acc - ref)) < 1e-9 * np.max(np.abs(ref))


def _oracle_align(archive, cases, template, fit_dm, niter):
    """ppalign.py:113-213 restated with oracle functions (exact FFTFIT polish)."""
    nsub, nchan, nbin = archive.nsub, archive.nchan, archive.nbin
    for _ in range(niter):
        aligned = np.zeros((nchan, nbin))
        tot = np.zeros(nchan)
        for s in range(nsub):
            ok = np.asarray(archive.ok_ichans[s])
            c = cases[s]
            port, model, freqs = c["data"][ok], template[ok], c["freqs"][ok]
            errs = archive.noise_stds[s, 0, ok]
            nu_fit = orc.guess_fit_freq(freqs, archive.SNRs[s, 0, ok])
            rot_port = orc.rotate_data(port, 0.0, archive.DM, c["P"], freqs, nu_fit)
            g = orc.fit_phase_shift(np.average(rot_port, axis=0, weights=archive.weights[s, ok]),
                                    model.mean(axis=0), Ns=nbin, polish="exact")
            r = orc.fit_portrait_full(port, model, [g.phase, archive.DM, 0, 0, 0], c["P"], freqs,
                                      [nu_fit] * 3, [None] * 3, errs, [1, int(fit_dm), 0, 0, 0],
                                      log10_tau=False)
            w = r.scales / errs ** 2
            aligned[ok] += w[:, None] * orc.rotate_data(port, r.phi, r.DM, c["P"], freqs, r.nu_DM)
            tot[ok] += w
        good = tot > 0
        aligned[good] /= tot[good, None]
        template = aligned
    return template, tot


def test_align_archives_two_iterations():
    """Config-5 style loop at a size the oracle finishes in seconds."""
    from pulseportraiture_b200 import ppalign
    archive, cases = _fake_archive(6, 32, 256, 9200, DM_stored=0.0)
    # the initial template is a noisy, slightly wrong model (a smoothed single subint)
    tmpl0 = orc.rotate_data(cases[0]["model"], 0.013)
    out = ppalign.align_archives([archive], tmpl0, fit_dm=True, niter=2, quiet=True)
    ref, tot = _oracle_align(archive, cases, tmpl0, True, 2)
    assert rel(out.weights, tot) < 1e-4
    assert np.max(np.abs(out.port - ref)) < 2e-5 * np.max(np.abs(ref))
    # 'place' and 'norm' options run and do what they say
    # (the reference's 1e-4-wide delta is only non-zero when `place` is within 20 sigma of a bin
    # centre, pplib.py:805-806: use a bin centre)
    place = 128.5 / 256.0
    out2 = ppalign.align_archives([archive], tmpl0, niter=1, norm="max", place=place, quiet=True)
    assert np.allclose(out2.port.max(axis=1)[out2.weights > 0], 1.0, atol=0.15)   # rotated after norm
    assert abs((np.argmax(out2.port.mean(axis=0)) + 0.5) / 256.0 - place) < 0.02


def test_zap_scans():
    from pulseportraiture_b200 import pptoas, ppzap
    archive, cases = _fake_archive(3, 32, 512, 9300)
    # RFI: channel 7 of subint 1 gets extra noise after the noise levels were measured
    rng = np.random.RandomState(1)
    archive.subints[1, 0, 7] += rng.normal(0, 3.0, 512)
    for s in range(3):
        archive.ok_ichans[s] = np.arange(32)
    gt = pptoas.GetTOAs([archive], synth.GMODEL, quiet=True)
    gt.get_TOAs(bary=False)
    gt.get_channels_to_zap(SNR_threshold=0.0, rchi2_threshold=2.0)
    assert gt.zap_channels[0][1] == [7] and gt.zap_channels[0][0] == [] and gt.zap_channels[0][2] == []
    # per-channel reduced chi2 vs the time-domain restatement (pptoas.py:1398-1401, pplib.py:727-750)
    s = 1
    r = gt
    port = orc.rotate_data(archive.subints[s, 0], r.phis[0][s], r.DMs[0][s], archive.Ps[s],
                           cases[s]["freqs"], r.nu_refs[0][s][0])
    model = cases[s]["model"]      # float32-rounded model, as the device saw it
    red = np.sum(((port - r.scales[0][s][:, None] * model) / archive.noise_stds[s, 0][:, None]) ** 2,
                 axis=1) / 510.
    assert rel(np.array(gt.channel_red_chi2s[0][s]), red) < 1e-4
    # median / sigma noise zapping (ppzap.py:18-48)
    archive.noise_stds[2, 0, [3, 20]] *= 8.0
    z = ppzap.get_zap_channels(archive, nstd=3)
    assert z[2] == [3, 20] and z[0] == []
    # S/N cut with iteration
    gt2 = pptoas.GetTOAs([archive], synth.GMODEL, quiet=True)
    gt2.get_TOAs(bary=False)
    srt = np.sort(gt2.channel_snrs[0][0])
    thr = np.sqrt(32) * 0.5 * (srt[2] + srt[3])
    gt2.get_channels_to_zap(SNR_threshold=thr, rchi2_threshold=1e9, iterate=False)
    assert len(gt2.zap_channels[0][0]) == 3


@pytest.mark.parametrize("nchan,nbin,nsub,chunk", [(24, 2048, 9, 4), (40, 512, 6, 0), (5, 64, 3, 0)])
def test_fused_fit_and_align_matches_two_steps(nchan, nbin, nsub, chunk):
    """pp_fit_batch(align_sum) = fit, then pp_align_accumulate with the fitted phi, DM, nu_out and
    w = scales / sigma^2 (ppalign.py:197-208); also against the oracle's rotate_data."""
    from pulseportraiture_b200.engine import WidebandPlan
    sigma = 1.5 if nbin >= 512 else 0.4
    cases = [synth.make_case(nchan, nbin, 1500., 800., 9300 + 7 * nbin + s, sigma=sigma) for s in range(nsub)]
    data = np.stack([c["data"] for c in cases]).astype(np.float32)
    P, freqs = cases[0]["P"], cases[0]["freqs"]
    mask = np.ones((nsub, nchan), dtype=np.uint8)
    mask[1, 2] = 0
    mask[nsub - 1, :nchan // 2] = 0
    with WidebandPlan(nchan, nbin) as pl:
        pl.set_model(cases[0]["model"].astype(np.float32), freqs)
        if chunk:
            pl.set_chunk(chunk)
        r = pl.fit_batch(data, P, chan_mask=mask, nu_fit_mode=1, Ns=nbin, align=True)
        w = np.where(mask > 0, r["scales"] / np.where(r["noise"] > 0, r["noise"], 1.0) ** 2, 0.0)
        acc, wsum = pl.align_accumulate(data, r["params"][:, 0], r["params"][:, 1], P, r["nu_out"][:, 0], w)
    assert rel(r["align_wsum"], wsum) < 1e-13
    # the fused path accumulates float32-rounded spectra: 6e-8 per term
    scale = np.max(np.abs(acc))
    assert np.max(np.abs(r["align_sum"] - acc)) < 3e-7 * scale
    ref = np.zeros((nchan, nbin))
    for s in range(nsub):
        ref += w[s][:, None] * orc.rotate_data(data[s].astype(np.float64), r["params"][s, 0], r["params"][s, 1], P,
                                               freqs, r["nu_out"][s, 0])
    assert np.max(np.abs(r["align_sum"] - ref)) < 3e-7 * scale


def test_get_narrowband_TOAs_matches_per_channel_fftfit():
    """GetTOAs.get_narrowband_TOAs (pptoas.py:745-1132): one 1-D FFTFIT per usable channel, all in
    one pp_fit_phase_shift_batch call with nmodel = nchan."""
    from pulseportraiture_b200 import pptoas
    d, cases = _fake_archive(4, 24, 512, 9500, sigma=0.5)
    gt = pptoas.GetTOAs([d], cases[0]["model"], quiet=True)
    gt.get_narrowband_TOAs(print_phase=True)
    nsub, nchan = 4, 24
    assert gt.phis[0].shape == (nsub, nchan)
    ntoa = sum(len(d.ok_ichans[s]) for s in range(nsub))
    assert len(gt.TOA_list) == ntoa
    rng = np.random.RandomState(0)
    for s in range(nsub):
        okc = np.asarray(d.ok_ichans[s])
        bad = np.setdiff1d(np.arange(nchan), okc)
        assert np.all(gt.phis[0][s, bad] == 0) and np.all(gt.scales[0][s, bad] == 0)
        for ch in rng.choice(okc, size=5, replace=False):
            noise = d.noise_stds[s, 0, ch]
            o = orc.fit_phase_shift(cases[s]["data"][ch], cases[0]["model"][ch], noise=noise, Ns=100, polish="exact")
            assert abs(gt.phis[0][s, ch] - o.phase) < SIG_TOL * o.phase_err
            assert rel(gt.phi_errs[0][s, ch], o.phase_err) < 1e-6
            assert rel(gt.scales[0][s, ch], o.scale) < 1e-6
            assert rel(gt.channel_snrs[0][s, ch], o.snr) < 1e-6
            assert rel(gt.channel_red_chi2s[0][s, ch], o.red_chi2) < 1e-7
            P = d.Ps[s]
            want = d.epochs[s].in_days() + (gt.phis[0][s, ch] * P + d.backend_delay) / 86400.0
            assert abs(gt.TOAs[0][s, ch].in_days() - want) < 1e-9
            assert rel(gt.TOA_errs[0][s, ch], gt.phi_errs[0][s, ch] * P * 1e6) < 1e-14
    t0 = gt.TOA_list[0]
    assert t0.flags["chan"] == int(d.ok_ichans[0][0]) and t0.flags["subint"] == 0 and "phs" in t0.flags
    assert t0.frequency == d.freqs[0, d.ok_ichans[0][0]]


def test_align_archives_on_a_different_frequency_grid():
    """An archive whose channels are a shifted sub-band of the template archive's grid: every data
    channel is fit against, and added to, the nearest template channel (ppalign.py:166-176)."""
    from pulseportraiture_b200 import ppalign
    from pulseportraiture_b200.pplib import DataBunch
    full, cases = _fake_archive(5, 32, 256, 9700)
    sel = np.arange(8, 24)
    sub = DataBunch(**dict(full))
    sub["nchan"] = len(sel)
    sub["freqs"] = full.freqs[:, sel] + 0.3                     # 0.3 MHz off the template's centres
    sub["subints"] = full.subints[:, :, sel]
    sub["noise_stds"] = full.noise_stds[:, :, sel]
    sub["SNRs"] = full.SNRs[:, :, sel]
    sub["weights"] = full.weights[:, sel]
    sub["ok_ichans"] = [np.arange(len(sel)) for _ in range(5)]
    tmpl = DataBunch(**dict(full))                              # template archive: first subint, full grid
    tmpl["subints"] = np.asarray(orc.rotate_data(cases[0]["model"], 0.013))[None, None]
    tmpl["nsub"] = 1
    out = ppalign.align_archives([sub], tmpl, fit_dm=True, niter=2, quiet=True)
    assert out.port.shape == (32, 256)
    outside = np.setdiff1d(np.arange(32), sel)
    assert np.all(out.weights[outside] == 0) and np.all(out.weights[sel] > 0)
    # the same archive aligned against the matching template rows given as a plain array
    ref = ppalign.align_archives([sub], np.asarray(tmpl.subints[0, 0])[sel], fit_dm=True, niter=1, quiet=True)
    one = ppalign.align_archives([sub], tmpl, fit_dm=True, niter=1, quiet=True)
    assert np.array_equal(one.port[sel], ref.port) and np.array_equal(one.weights[sel], ref.weights)
    # several data channels per template channel: the reference keeps only the last one (numpy +=)
    dup = DataBunch(**dict(sub))
    dup["freqs"] = np.tile(np.linspace(1400., 1410., len(sel)), (5, 1))
    many = ppalign.align_archives([dup], tmpl, niter=1, quiet=True)      # tests/test_gpu_round2.py checks the values
    assert many.port.shape == (32, 256) and 1 <= np.count_nonzero(many.weights) < len(sel)
