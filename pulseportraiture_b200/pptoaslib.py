"""Reference-compatible facade for ``pptoaslib.fit_portrait_full``
(pptoaslib.py:928-1096) on top of the batched C ABI."""
from __future__ import annotations

import sys
import time

import numpy as np

from .pplib import (DataBunch, RCSTRINGS, Dconst, _check_bounds, _f32, _dev, _mdl,  # noqa: F401
                    get_plan, scattering_times, scattering_portrait_FT, scipy_return_code, _RC_BENIGN)


def gaussian_profile_FT(nbin, loc, wid, amp):
    """Analytic Fourier transform of a Gaussian profile of FWHM ``wid`` [rot] at phase ``loc``,
    windowed (Gaussian convolved with a sinc), sampled at the nbin/2 + 1 harmonics
    (pptoaslib.py:14-50).  Host arithmetic: a few kB per call."""
    from scipy.special import erf
    nharm = nbin // 2 + 1
    if wid <= 0.0:
        return np.zeros(nharm, 'd')
    sigma = wid / (2 * np.sqrt(2 * np.log(2)))
    amp = amp * (2 * np.pi * sigma ** 2) ** 0.5
    sigma = 1.0 / (sigma * 2 * np.pi)
    harmind = np.arange(nharm)
    a = sigma / ((1.0 / np.pi) * 2 ** 0.5)
    b = harmind / (sigma * 2 ** 0.5)
    retvals = np.exp(-b ** 2) * (erf(a - b * 1j) + erf(a + b * 1j)) / 2
    retvals = retvals * (amp * nbin)
    if loc != 0.0:
        retvals = retvals * np.exp(-harmind * 2.0j * np.pi * loc)
    return np.nan_to_num(retvals)


def instrumental_response_FT(nbin, wid=0.0, irf_type='rect'):
    """Fourier transform of an instrumental response of width ``wid`` [rot]: a rectangle ('rect')
    or a Gaussian of that FWHM ('gauss') (pptoaslib.py:112-145)."""
    nharm = nbin // 2 + 1
    if wid == 0.0:
        return np.ones(nharm)
    if irf_type == 'rect':
        return np.sinc(np.arange(nharm) * wid)
    if irf_type == 'gauss':
        gp_FT = gaussian_profile_FT(nbin, 0.0, wid, 1.0)
        return gp_FT / gp_FT[0]
    print("Unrecognized instrumental response function type '%s'." % irf_type)
    return 0


def instrumental_response_port_FT(nbin, freqs, DM=0.0, P=1.0, wids=[], irf_types=[], chan_bw=None):
    """Combined instrumental responses per channel and harmonic, [nchan, nbin/2 + 1]
    (pptoaslib.py:147-179): the product of the constant responses ``wids`` / ``irf_types`` and,
    when DM is non-zero, of a rectangle of the per-channel smearing width
    8.3e-6 chan_bw / (nu/GHz)^3 / P (the expression of pptoaslib.py:175, which carries no DM factor).
    ``chan_bw`` (extra argument) overrides abs(freqs[1] - freqs[0]): get_TOAs passes the spacing of the
    first two *usable* channels, which is what the reference computes from ``freqsx``."""
    freqs = np.asarray(freqs, dtype=np.float64)
    nharm = nbin // 2 + 1
    nchan = len(freqs)
    if DM == 0.0 and len(wids) == 0:
        return np.ones([nchan, nharm])
    resp = np.ones([nchan, nharm], dtype=complex)
    for wid, irf_type in zip(wids, irf_types):
        resp = resp * instrumental_response_FT(nbin, wid, irf_type)[None, :]
    if DM:
        if chan_bw is None:
            chan_bw = abs(freqs[1] - freqs[0])
        for ichan, freq in enumerate(freqs):
            wid = 8.3e-6 * chan_bw / (freq / 1e3) ** 3 / P
            resp[ichan] = resp[ichan] * instrumental_response_FT(nbin, wid, 'rect')
    return resp


def add_instrumental_response(model_port, freqs, DM=0.0, P=1.0, wids=[], irf_types=[], chan_bw=None):
    """model -> irfft(instrumental_response_port_FT(...) * rfft(model)) (pptoas.py:388-394); the
    per-harmonic multiply runs on the device (pp_apply_response_batch).  The responses of
    pptoaslib.py:112-145 with loc = 0 are real."""
    model_port = np.asarray(model_port)
    nchan, nbin = model_port.shape
    resp = instrumental_response_port_FT(nbin, freqs, DM, P, wids, irf_types, chan_bw=chan_bw)
    if np.iscomplexobj(resp):
        if np.abs(resp.imag).max() > 1e-14 * np.abs(resp.real).max():
            raise ValueError("complex instrumental response")
        resp = resp.real
    pl = get_plan(nchan, nbin)
    return pl.apply_response_batch(_f32(model_port)[None], np.ascontiguousarray(resp, dtype=np.float64))[0] \
        .astype(np.float64)


def fit_portrait_full(data_port, model_port, init_params, P, freqs,
                      nu_fits=[None, None, None], nu_outs=[None, None, None],
                      errs=None, fit_flags=[1, 1, 1, 1, 1],
                      bounds=[(None, None), (None, None), (None, None),
                              (None, None), (None, None)], log10_tau=True,
                      option=0, sub_id=None, method='trust-ncg', is_toa=True,
                      quiet=True):
    """Fit phase, DM, GM, tau and alpha between data and model portraits.

    Same arguments, units and DataBunch fields as pptoaslib.fit_portrait_full
    (pptoaslib.py:928-1096).  ``method`` is accepted for compatibility; the
    scipy minimisers are replaced by the on-device safeguarded Newton solver,
    which converges to the same optimum; ``return_code`` is the scipy status ``method`` reports
    for the same outcome (trust-ncg: 2 normal exit, 1 pass limit, 3 non-finite; pplib._RC_MAP), the
    device's own code is the extra field ``device_return_code``.  ``bounds`` are used with
    method='TNC' only, as in the reference (pptoaslib.py:1008-1014); the
    solver treats them as an active set.
    """
    if method == 'TNC':
        bounds = _check_bounds(bounds, 5)
    elif method not in ('trust-ncg', 'Newton-CG'):
        print("Method '%s' is not implemented." % method)
        sys.exit()
    data_port = np.asarray(data_port)
    nchan, nbin = data_port.shape
    freqs = np.asarray(freqs, dtype=np.float64)
    pl = get_plan(nchan, nbin)
    pl.set_model(_mdl(model_port), freqs)
    init = np.array(init_params, dtype=np.float64).reshape(1, 5)

    def three(vals):
        vals = list(vals)
        if all(v is None for v in vals):
            return None
        return np.array([[np.nan if v is None else float(v) for v in vals]])

    start = time.time()
    r = pl.fit_batch(_dev(data_port)[None], P,
                     errs=None if errs is None else np.asarray(errs, dtype=np.float64)[None],
                     init=init, nu_fits=three(nu_fits), nu_outs=three(nu_outs),
                     fit_flags=[1 if f else 0 for f in fit_flags],
                     log10_tau=bool(log10_tau), option=int(option),
                     is_toa=bool(is_toa), semantics="full",
                     bounds=bounds if method == 'TNC' else None)
    duration = time.time() - start
    drc = int(r["return_code"][0])
    rc = scipy_return_code(drc, method)
    if rc not in _RC_BENIGN[method]:
        if sub_id is not None:
            ii = sub_id[::-1].index("_")
            isub, filename = sub_id[-ii:], sub_id[:-ii - 1]
            sys.stderr.write("Fit 'failed' with return code %d: %s -- %s subint %s\n"
                             % (rc, RCSTRINGS[str(rc)], filename, isub))
        else:
            sys.stderr.write("Fit 'failed' with return code %d -- %s" % (rc, RCSTRINGS[str(rc)]))
    ifit = np.where(fit_flags)[0]
    params = list(r["params"][0])
    perr = r["param_errs"][0].copy()
    cov = r["cov"][0][np.ix_(ifit, ifit)]
    return DataBunch(params=params, param_errs=perr, phi=params[0],
                     phi_err=perr[0], DM=params[1], DM_err=perr[1],
                     GM=params[2], GM_err=perr[2], tau=params[3],
                     tau_err=perr[3], alpha=params[4], alpha_err=perr[4],
                     scales=r["scales"][0], scale_errs=r["scale_errs"][0],
                     nu_DM=r["nu_out"][0, 0], nu_GM=r["nu_out"][0, 1],
                     nu_tau=r["nu_out"][0, 2], covariance_matrix=cov,
                     chi2=r["chi2"][0], red_chi2=r["red_chi2"][0],
                     snr=r["snr"][0], channel_snrs=r["channel_snrs"][0],
                     duration=duration, nfeval=int(r["nfeval"][0]),
                     return_code=rc, device_return_code=drc)


def rotate_portrait_full(port, phi, DM, GM, freqs, nu_DM=np.inf, nu_GM=np.inf, P=None):
    """Rotate / dedisperse a portrait including the nu**-4 term (pptoaslib.py:52-81)."""
    port = np.asarray(port)
    nchan, nbin = port.shape
    if P is None:
        P = 1.0
    big = lambda v: 1e300 if np.isinf(v) else float(v)  # noqa: E731
    pl = get_plan(nchan, nbin)
    pl.set_freqs(np.asarray(freqs, dtype=np.float64))
    out = pl.rotate_batch(_f32(port)[None], phi, DM, P, big(nu_DM), GM=GM, nu_GM=big(nu_GM))
    return out[0].astype(np.float64)


def get_scales_full(params, data_port, model_port, P, freqs, nu_DM, nu_GM, nu_tau, log10_tau,
                    errs=None):
    """Maximum-likelihood per-channel amplitudes at given parameters
    (pptoaslib.py:908-926; takes time-domain portraits instead of their FFTs)."""
    data_port = np.asarray(data_port)
    nchan, nbin = data_port.shape
    pl = get_plan(nchan, nbin)
    pl.set_model(_mdl(model_port), np.asarray(freqs, dtype=np.float64))
    nus = np.array([[nu_DM, nu_GM, nu_tau]], dtype=np.float64)
    r = pl.fit_batch(_f32(data_port)[None], P,
                     errs=None if errs is None else np.asarray(errs, dtype=np.float64)[None],
                     init=np.array(params, dtype=np.float64).reshape(1, 5), nu_fits=nus, nu_outs=nus,
                     fit_flags=(1, 1, 1, 1, 1), log10_tau=bool(log10_tau), max_iter=-1)
    return r["scales"][0]
