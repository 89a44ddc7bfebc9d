"""Reference-compatible facade for the hot-path functions of ``pplib``.

Same names, argument meaning, units and return fields as the reference
(pennucci/PulsePortraiture ``pplib.py``; line numbers cited per function), but
the arithmetic runs in the sm_100a CUDA library behind the C ABI
(``include/ppb200.h``).  Inputs are converted to float32 on entry (the device
storage type); parameters and accumulators are float64.

Not accelerated (plain host helpers kept so that scripts run): model
generation from ``.gmodel`` files and scalar reference-frequency transforms.
"""
from __future__ import annotations

import sys
import time

import numpy as np

from .engine import WidebandPlan

# ---- settings (pplib.py:44-83) ---------------------------------------------
Dconst_exact = 4.148808e3
Dconst_trad = 0.000241 ** -1
Dconst = Dconst_trad
scattering_alpha = -4.0
use_get_noise = True
default_noise_method = "PS"
F0_fact = 0
wid_max = 0.25
default_model = "000"
binshift = 1.0

# Return codes.  The device solver reports 0 converged / 1 pass limit / 3 non-finite objective
# (DEVICE_RCSTRINGS).  The facades hand callers the scipy status the reference's minimiser would
# have returned for the same outcome (pplib.py:2159-2171, pptoaslib.py:1018-1033: callers treat
# TNC's {1, 2, 4} and trust-ncg's 2 as the normal exits), with the messages of pplib.py:111-119;
# the device's own code stays available as the extra DataBunch field ``device_return_code``.
DEVICE_RCSTRINGS = {"0": "CONVERGED: Newton step below tolerance.",
                    "1": "MAXITER: Maximum number of objective passes reached.",
                    "3": "NONFINITE: Objective is not finite."}
RCSTRINGS = {'-1': 'INFEASIBLE: Infeasible (low > up).',
             '0': 'LOCALMINIMUM: Local minima reach (|pg| ~= 0).',
             '1': 'FCONVERGED: Converged (|f_n-f_(n-1)| ~= 0.)',
             '2': 'XCONVERGED: Converged (|x_n-x_(n-1)| ~= 0.)',
             '3': 'MAXFUN: Max. number of function evaluations reach.',
             '4': 'LSFAIL: Linear search failed.',
             '5': 'CONSTANT: All lower bounds are equal to the upper bounds.',
             '6': 'NOPROGRESS: Unable to progress.',
             '7': 'USERABORT: User requested end of minimization.'}
# device code -> scipy status, per minimiser
_RC_MAP = {'TNC': {0: 1, 1: 3, 3: 6},            # FCONVERGED / MAXFUN / NOPROGRESS
           'trust-ncg': {0: 2, 1: 1, 3: 3},      # "failure to predict improvement" is the normal exit
           'Newton-CG': {0: 0, 1: 1, 3: 3}}      # success / maxiter / NaN result encountered
_RC_BENIGN = {'TNC': (0, 1, 2, 4), 'trust-ncg': (0, 1, 2, 4), 'Newton-CG': (0, 1, 2, 4)}


def scipy_return_code(device_rc, method='TNC'):
    """The status scipy's ``method`` reports for the outcome the device solver reports as ``device_rc``
    (scalar or array)."""
    m = _RC_MAP[method]
    if np.ndim(device_rc) == 0:
        return m.get(int(device_rc), int(device_rc))
    return np.array([m.get(int(v), int(v)) for v in np.asarray(device_rc).ravel()],
                    dtype=int).reshape(np.shape(device_rc))


class DataBunch(dict):
    """dict with attribute access (pplib.py:125-136)."""

    def __init__(self, **kwds):
        dict.__init__(self, kwds)
        self.__dict__ = self


# ---- plan cache ---------------------------------------------------------------
_plans = {}
_device = 0


def set_device(device):
    """Select the CUDA device used by the single-portrait facade calls."""
    global _device
    _device = int(device)


def get_plan(nchan, nbin, device=None):
    device = _device if device is None else int(device)
    key = (int(nchan), int(nbin), device)
    pl = _plans.get(key)
    if pl is None:
        pl = WidebandPlan(nchan, nbin, device)
        _plans[key] = pl
    return pl


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _mdl(a):
    """A model portrait as set_model takes it: float32 arrays as they are, everything else as float64 (the
    reference's array type; no float32 rounding floor in the model spectrum)."""
    a = np.asarray(a)
    return np.ascontiguousarray(a, dtype=np.float32 if a.dtype == np.float32 else np.float64)


def _dev(a):
    """Portrait data as the device takes them without a host pass: float32 and float64 arrays go as
    they are (float64 is rounded to float32 on the device, PP_DATA_F64), anything else via float64."""
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def _check_bounds(bounds, nparam):
    """scipy-style bounds -> what WidebandPlan.fit_batch takes (None when nothing is bounded)."""
    bounds = list(bounds or [])
    if len(bounds) > nparam:
        raise ValueError("bounds has %d entries for %d parameters" % (len(bounds), nparam))
    for b in bounds:
        if b is not None and len(b) != 2:
            raise ValueError("each bound is a (lower, upper) pair")
    return bounds if any(b is not None and any(v is not None for v in b) for b in bounds) else None


# ---- A3: 1-D FFTFIT -------------------------------------------------------------
def fit_phase_shift(data, model, noise=None, bounds=[-0.5, 0.5], Ns=100):
    """Fit a phase shift between data and model (pplib.py:2054-2100).

    Returns DataBunch(phase, phase_err, scale, scale_err, snr, red_chi2,
    duration).  The brute-force grid (Ns points on ``bounds``, both ends) is
    evaluated on the device; ``phase`` is the exact minimiser reached from the
    grid argmin (the reference's Nelder-Mead polish is accurate to ~1e-4 rot).
    The integer argmin is returned as the extra field ``lag_index``.
    """
    data = np.asarray(data)
    if len(bounds) != 2 or not bounds[1] > bounds[0]:
        raise ValueError("bounds = [lower, upper]")
    nbin = data.shape[-1]
    pl = get_plan(1, nbin)
    start = time.time()
    r = pl.fit_phase_shift_batch(_f32(data).reshape(1, nbin),
                                 _f32(model).reshape(1, nbin),
                                 None if noise is None else np.array([noise], dtype=np.float64),
                                 Ns=Ns, bounds=bounds)
    duration = time.time() - start
    return DataBunch(phase=r["phase"][0], phase_err=r["phase_err"][0],
                     scale=r["scale"][0], scale_err=r["scale_err"][0],
                     snr=r["snr"][0], red_chi2=r["red_chi2"][0],
                     duration=duration, lag_index=int(r["lag_index"][0]))


# ---- A1: phi + DM fit --------------------------------------------------------------
def fit_portrait(data, model, init_params, P, freqs, nu_fit=None, nu_out=None,
                 errs=None, bounds=[(None, None), (None, None)], id=None,
                 quiet=True):
    """Fit a phase offset and DM between data and model portraits
    (pplib.py:2102-2204).  Same arguments and DataBunch fields as the
    reference; the TNC minimiser is replaced by the on-device Newton solver
    (``return_code`` 0 = converged, 1 = max passes, 3 = non-finite), which
    honours ``bounds`` as an active set."""
    bounds = _check_bounds(bounds, 2)
    data = np.asarray(data)
    nchan, nbin = data.shape
    freqs = np.asarray(freqs, dtype=np.float64)
    pl = get_plan(nchan, nbin)
    pl.set_model(_mdl(model), freqs)
    init = np.zeros((1, 5))
    init[0, 0], init[0, 1] = init_params[0], init_params[1]
    nu_fits = None if nu_fit is None else np.full((1, 3), float(nu_fit))
    nu_outs = None if nu_out is None else np.full((1, 3), float(nu_out))
    start = time.time()
    r = pl.fit_batch(_dev(data)[None], P,
                     errs=None if errs is None else np.asarray(errs, dtype=np.float64)[None],
                     init=init, nu_fits=nu_fits, nu_outs=nu_outs,
                     fit_flags=(1, 1, 0, 0, 0), semantics="fit_portrait", bounds=bounds)
    duration = time.time() - start
    drc = int(r["return_code"][0])
    rc = scipy_return_code(drc, 'TNC')
    if not quiet and rc not in _RC_BENIGN['TNC']:
        if id is not None:
            ii = id[::-1].index("_")
            isub, filename = id[-ii:], id[:-ii - 1]
            sys.stderr.write("Fit failed with return code %d: %s -- %s subint %s\n"
                             % (rc, RCSTRINGS[str(rc)], filename, isub))
        else:
            sys.stderr.write("Fit failed with return code %d -- %s" % (rc, RCSTRINGS[str(rc)]))
    return DataBunch(phase=r["params"][0, 0], phase_err=r["param_errs"][0, 0],
                     DM=r["params"][0, 1], DM_err=r["param_errs"][0, 1],
                     scales=r["scales"][0], scale_errs=r["scale_errs"][0],
                     nu_ref=r["nu_out"][0, 0], covariance=r["cov"][0, 0, 1],
                     chi2=r["chi2"][0], red_chi2=r["red_chi2"][0],
                     snr=r["snr"][0], duration=duration,
                     nfeval=int(r["nfeval"][0]), return_code=rc, device_return_code=drc)


def get_scales(data, model, phase, DM, P, freqs, nu_ref=np.inf):
    """Best-fit per-channel amplitudes at given (phase, DM) (pplib.py:2310-2336)."""
    data = np.asarray(data)
    nchan, nbin = data.shape
    pl = get_plan(nchan, nbin)
    pl.set_model(_mdl(model), np.asarray(freqs, dtype=np.float64))
    init = np.zeros((1, 5))
    init[0, 0], init[0, 1] = phase, DM
    nu = 1e300 if np.isinf(nu_ref) else float(nu_ref)
    r = pl.fit_batch(_f32(data)[None], P, init=init, nu_fits=np.full((1, 3), nu),
                     nu_outs=np.full((1, 3), nu), fit_flags=(1, 1, 0, 0, 0),
                     max_iter=-1, semantics="fit_portrait")
    return r["scales"][0]


# ---- A9: noise ------------------------------------------------------------------------
def get_noise(data, method=default_noise_method, **kwargs):
    """Off-pulse noise estimate (pplib.py:2206-2225): "PS" or "fit"."""
    if method == "PS":
        return get_noise_PS(data, **kwargs)
    if method == "fit":
        return get_noise_fit(data, **kwargs)
    print("Unknown get_noise method.")
    return 0


def get_noise_fit(data, fact=1.1, chans=False):
    """Noise from the harmonics above fact * the cutoff a fit of b exp(-a k) + dc to the log power
    spectrum finds (pplib.py:2255-2284, find_kc 1465-1495)."""
    data = np.asarray(data)
    if chans:
        nchan, nbin = data.shape
        return get_plan(nchan, nbin).get_noise_fit_batch(_f32(data)[None], fact)[0]
    rav = data.ravel()
    n = rav.size
    return get_plan(1, n).get_noise_fit_batch(_f32(rav).reshape(1, 1, n), fact)[0, 0]


def get_noise_PS(data, frac=4, chans=False):
    """Mean of the highest 1/frac of the power spectrum (pplib.py:2227-2253)."""
    data = np.asarray(data)
    if chans:
        nchan, nbin = data.shape
        kc = int((1 - frac ** -1) * (nbin // 2 + 1))         # pplib.py:2244
        return get_plan(nchan, nbin).get_noise_batch(_f32(data)[None], kc=kc)[0]
    rav = data.ravel()
    n = rav.size
    kc = int((1 - frac ** -1) * (n // 2 + 1))
    return get_plan(1, n).get_noise_batch(_f32(rav).reshape(1, 1, n), kc=kc)[0, 0]


# ---- A10: rotation ----------------------------------------------------------------------
def rotate_data(data, phase=0.0, DM=0.0, Ps=None, freqs=None, nu_ref=np.inf):
    """Rotate and/or dedisperse a profile, portrait or [nsub,npol,nchan,nbin]
    cube in the Fourier domain (pplib.py:2338-2426).  Positive phase/DM rotate
    to earlier phase."""
    data = np.asarray(data)
    shape = data.shape
    nbin = shape[-1]
    nu = 1e300 if np.isinf(nu_ref) else float(nu_ref)
    if data.ndim == 1:
        cube = data.reshape(1, 1, nbin)
        f = np.array([1.0 if freqs is None else float(np.atleast_1d(freqs)[0])])
        Pv = 1.0 if Ps is None else Ps
    elif data.ndim == 2:
        cube = data.reshape(1, shape[0], nbin)
        f = np.ones(shape[0]) if freqs is None else np.asarray(freqs, dtype=np.float64)
        Pv = 1.0 if Ps is None else Ps
    elif data.ndim == 4:
        nsub, npol, nchan, _ = shape
        cube = data.reshape(nsub * npol, nchan, nbin)
        f = np.asarray(freqs, dtype=np.float64)
        Pv = np.repeat(np.ones(nsub) * (1.0 if Ps is None else Ps), npol)
        if f.ndim == 2 and np.any(f != f[0]):
            # per-subint frequency arrays (pplib.py:2398-2411): one batch per distinct table
            if DM != 0.0 and Ps is None:
                raise ValueError("Ps and freqs are needed when DM != 0")
            tables, table_of = np.unique(f, axis=0, return_inverse=True)
            table_of = np.repeat(np.asarray(table_of).reshape(-1), npol)
            pl = get_plan(nchan, nbin)
            out = np.empty(cube.shape, dtype=np.float32)
            for t in range(len(tables)):
                sel = np.where(table_of == t)[0]
                pl.set_freqs(tables[t])
                out[sel] = pl.rotate_batch(_f32(cube[sel]), phase, DM if DM else 0.0, Pv[sel], nu)
            return out.astype(np.float64).reshape(shape)
        if f.ndim == 2:
            f = f[0]
    else:
        raise ValueError("Wrong number of dimensions.")
    if DM != 0.0 and (Ps is None or freqs is None):
        raise ValueError("Ps and freqs are needed when DM != 0")
    pl = get_plan(cube.shape[1], nbin)
    pl.set_freqs(f)
    out = pl.rotate_batch(_f32(cube), phase, DM if DM else 0.0, Pv, nu)
    return out.astype(np.float64).reshape(shape)


def rotate_portrait(port, phase=0.0, DM=None, P=None, freqs=None, nu_ref=np.inf):
    """pplib.py:2428-2460."""
    if DM is None and freqs is None:
        return rotate_data(port, phase)
    return rotate_data(port, phase, DM, P, freqs, nu_ref)


def rotate_profile(profile, phase=0.0):
    """pplib.py:2548-2559."""
    return rotate_data(profile, phase)


# ---- A11: scalar transforms (host arithmetic) --------------------------------------------
def DM_delay(DM, freq, freq_ref=np.inf, P=None):
    """pplib.py:2577-2590."""
    delay = Dconst * DM * ((freq ** -2.0) - (freq_ref ** -2.0))
    return delay / P if P else delay


def phase_transform(phi, DM, nu_ref1=np.inf, nu_ref2=np.inf, P=None, mod=False):
    """pplib.py:2592-2616."""
    if P is None:
        P, mod = 1.0, False
    phi_prime = phi + (Dconst * DM * P ** -1 * (nu_ref2 ** -2.0 - nu_ref1 ** -2.0))
    if mod:
        phi_prime = np.where(abs(phi_prime) >= 0.5, phi_prime % 1, phi_prime)
        phi_prime = np.where(phi_prime >= 0.5, phi_prime - 1.0, phi_prime)
        if not phi_prime.shape:
            phi_prime = np.float64(phi_prime)
    return phi_prime


def guess_fit_freq(freqs, SNRs=None):
    """pplib.py:2618-2632."""
    freqs = np.asarray(freqs, dtype=np.float64)
    nu0 = (freqs.min() + freqs.max()) * 0.5
    if SNRs is None:
        SNRs = np.ones(len(freqs))
    diff = np.sum((freqs - nu0) * SNRs * freqs ** -2) / np.sum(SNRs * freqs ** -2)
    return nu0 + diff


def guess_fit_freq_batch(freqs, SNRs, mask):
    """guess_fit_freq (pplib.py:2618-2632) of every row over its usable channels (mask != 0): [nsub, nchan] arrays
    in, [nsub] out; rows without usable channels give 0."""
    freqs = np.asarray(freqs, dtype=np.float64)
    m = np.asarray(mask) != 0
    w = np.where(m, np.asarray(SNRs, dtype=np.float64) * freqs ** -2, 0.0)
    den = w.sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        nu0 = 0.5 * (np.where(m, freqs, np.inf).min(axis=1) + np.where(m, freqs, -np.inf).max(axis=1))
        out = nu0 + ((freqs - nu0[:, None]) * w).sum(axis=1) / den
    return np.where(m.any(axis=1) & (den != 0), out, 0.0)


def scattering_times(tau, alpha, freqs, nu_tau):
    """pplib.py:4049-4053."""
    return tau * (np.asarray(freqs, dtype=np.float64) / nu_tau) ** alpha


def scattering_portrait_FT(taus, nbin, binshift=binshift):
    """1/(1 + 2 pi i k tau_n) (pplib.py:4080-4095)."""
    taus = np.atleast_1d(np.asarray(taus, dtype=np.float64))
    nharm = nbin // 2 + 1
    if not np.any(taus):
        return np.ones([len(taus), nharm])
    return 1.0 / (1.0 + 2.0j * np.pi * np.outer(taus, np.arange(nharm)))


# ---- synthetic model portraits (host setup code, not on the hot path) ----------------------
def get_bin_centers(nbin, lo=0.0, hi=1.0):
    """pplib.py:671-684."""
    lo, hi = np.double(lo), np.double(hi)
    diff = hi - lo
    return np.linspace(lo + diff / (nbin * 2), hi - diff / (nbin * 2), nbin)


def _wrapped_gaussian(nbin, loc, wid):
    """Unit-amplitude wrapped Gaussian of FWHM wid at phase loc
    (gaussian_profile, pplib.py:770-825, norm=False)."""
    out = np.zeros(nbin)
    if not wid > 0.0:
        return out
    sigma = wid / (2 * np.sqrt(2 * np.log(2)))
    mean = loc % 1.0
    x = get_bin_centers(nbin)
    x = np.where(x > mean + 0.5, x - 1.0, x) if mean < 0.5 else \
        np.where(x < mean - 0.5, x + 1.0, x)
    z = (x - mean) / sigma
    near = np.fabs(z) < 20.0
    out[near] = np.exp(-0.5 * z[near] ** 2) / (sigma * np.sqrt(2 * np.pi))
    if np.max(np.abs(out)) == 0.0:
        return out
    ipk = out.argmax()
    return out * (np.exp(-0.5 * ((x[ipk] - loc) / sigma) ** 2) / out[ipk])


def evolve_parameter(freqs, nu_ref, parameter, evol_parameter, code):
    """Power-law ('0') or linear ('1') evolution (pplib.py:996-1046)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    if code == "0":
        return np.exp(np.outer(np.log(freqs) - np.log(nu_ref), evol_parameter) +
                      np.outer(np.ones(len(freqs)), np.log(parameter)))
    if code == "1":
        return np.outer(freqs - nu_ref, evol_parameter) + \
            np.outer(np.ones(len(freqs)), parameter)
    raise KeyError(code)


def gen_gaussian_portrait(model_code, params, scattering_index, phases, freqs,
                          nu_ref, join_ichans=[], P=None):
    """Evolving-Gaussian model portrait (pplib.py:853-930).  params =
    [DC, tau_bin, (loc, m_loc, wid, m_wid, amp, m_amp) * ngauss]."""
    if len(join_ichans):
        raise NotImplementedError("join parameters are a ppgauss feature")
    params = np.asarray(params, dtype=np.float64)
    freqs = np.asarray(freqs, dtype=np.float64)
    nbin, nchan = len(phases), len(freqs)
    locs = evolve_parameter(freqs, nu_ref, params[2::6], params[3::6], model_code[0])
    wids = evolve_parameter(freqs, nu_ref, params[4::6], params[5::6], model_code[1])
    amps = evolve_parameter(freqs, nu_ref, params[6::6], params[7::6], model_code[2])
    gport = np.empty([nchan, nbin])
    for ichan in range(nchan):
        prof = np.zeros(nbin) + params[0]
        for ig in range(locs.shape[1]):
            prof += amps[ichan, ig] * _wrapped_gaussian(nbin, locs[ichan, ig], wids[ichan, ig])
        gport[ichan] = prof
    tau = params[1]
    if tau != 0.0:
        taus = scattering_times(float(tau) / nbin, scattering_index, freqs, nu_ref)
        gport = np.fft.irfft(scattering_portrait_FT(taus, nbin) *
                             np.fft.rfft(gport, axis=-1), axis=-1)
    return gport


def gen_gaussian_portrait_device(model_code, params, scattering_index, phases, freqs,
                                 nu_ref, join_ichans=[], P=None):
    """gen_gaussian_portrait (pplib.py:853-930) evaluated by the CUDA kernel
    ``k_gauss_model`` (+ the scattering multiply of ``k_rotate``): the per-archive /
    per-subint model build of pptoas.py:356-379.  ``phases`` must be the bin centres
    of get_bin_centers(nbin).  Returns float64 [nchan, nbin], evaluated in double (a scattered model,
    tau != 0, passes through float32 rows)."""
    if len(join_ichans):
        raise NotImplementedError("join parameters are a ppgauss feature")
    nbin, nchan = len(phases), len(freqs)
    if not np.allclose(phases, get_bin_centers(nbin), rtol=0, atol=1e-12):
        raise ValueError("the device generator works on get_bin_centers(nbin)")
    pl = get_plan(nchan, nbin)
    pl.set_freqs(np.asarray(freqs, dtype=np.float64))
    return pl.gen_gaussian_portrait(model_code, params, scattering_index, nu_ref, dtype=np.float64)


def gen_spline_portrait(mean_prof, freqs, eigvec, tck, nbin=None, device=False):
    """Model portrait from a make_spline_model(...) PCA / B-spline model
    (pplib.py:932-956): mean_prof + splev(freqs, tck).T . eigvec.T.  ``device=True``
    evaluates it with the CUDA kernel ``k_spline_model`` (float32 precision)."""
    mean_prof = np.asarray(mean_prof, dtype=np.float64)
    freqs = np.asarray(freqs, dtype=np.float64)
    eigvec = np.asarray(eigvec, dtype=np.float64).reshape(len(mean_prof), -1)
    shift = 0.0
    if nbin is not None and nbin != len(mean_prof):
        # pplib.py:951-955: Fourier resampling, then the half-bin-difference rotation that undoes
        # the shift ss.resample introduces.  Resampling is linear along the bin axis, so it is
        # applied to the basis (mean profile, eigenvectors) once instead of to every channel.
        import scipy.signal as ss
        shift = 0.5 * (nbin ** -1 - len(mean_prof) ** -1)
        mean_prof = ss.resample(mean_prof, nbin)
        eigvec = ss.resample(eigvec, nbin, axis=0)
    if device:
        pl = get_plan(len(freqs), len(mean_prof))
        pl.set_freqs(freqs)
        port = pl.gen_spline_portrait(mean_prof, eigvec, tck).astype(np.float64)
    elif not eigvec.shape[1]:
        port = np.tile(mean_prof, len(freqs)).reshape(len(freqs), len(mean_prof))
    else:
        import scipy.interpolate as si
        proj_port = np.array(si.splev(freqs, tck, der=0, ext=0)).T
        port = np.dot(proj_port, eigvec.T) + mean_prof
    if shift:
        port = rotate_portrait(port, shift)
    return port


def read_spline_model(modelfile, freqs=None, nbin=None, quiet=False, device=False):
    """Read a make_spline_model(...) pickle (pplib.py:2955-2987): without freqs returns
    (modelname, source, datafile, mean_prof, eigvec, tck), else (modelname, model)."""
    import pickle
    if not quiet:
        print("Reading model from %s..." % modelfile)
    with open(modelfile, "rb") as fh:
        modelname, source, datafile, mean_prof, eigvec, tck = pickle.load(fh, encoding="latin1")
    if freqs is None:
        return (modelname, source, datafile, mean_prof, eigvec, tck)
    return (modelname, gen_spline_portrait(mean_prof, freqs, eigvec, tck, nbin, device=device))


def is_spline_model(modelfile):
    """True if ``modelfile`` is a pickled spline model rather than a text .gmodel file
    (the reference falls back to read_spline_model when read_model fails, pptoas.py:376-379)."""
    import pickle
    try:
        with open(modelfile, "rb") as fh:
            obj = pickle.load(fh, encoding="latin1")
        return isinstance(obj, (list, tuple)) and len(obj) == 6
    except Exception:
        return False


def read_model(modelfile, phases=None, freqs=None, P=None, quiet=False, device=False):
    """Read a ``.gmodel`` file (pplib.py:2867-2953).  Without phases/freqs
    returns (name, code, nu_ref, ngauss, params, fit_flags, alpha, fit_alpha);
    otherwise (name, ngauss, model).  ``device=True`` builds the portrait on the GPU."""
    read_only = phases is None and freqs is None
    comps = []
    modelname, model_code, nu_ref = "", default_model, None
    dc = tau = 0.0
    alpha = scattering_alpha
    fit_dc = fit_tau = fit_alpha = 0
    with open(modelfile, "r") as fh:
        for line in fh:
            info = line.split()
            if not info:
                continue
            try:
                if info[0] == "MODEL":
                    modelname = info[1]
                elif info[0] == "CODE":
                    model_code = info[1]
                elif info[0] == "FREQ":
                    nu_ref = np.float64(info[1])
                elif info[0] == "DC":
                    dc, fit_dc = np.float64(info[1]), int(info[2])
                elif info[0] == "TAU":
                    tau, fit_tau = np.float64(info[1]), int(info[2])
                elif info[0] == "ALPHA":
                    alpha, fit_alpha = np.float64(info[1]), int(info[2])
                elif info[0][:4] == "COMP":
                    comps.append(info)
            except IndexError:
                pass
    ngauss = len(comps)
    params = np.zeros(ngauss * 6 + 2)
    fit_flags = np.zeros(len(params))
    params[0], params[1] = dc, tau
    fit_flags[0], fit_flags[1] = fit_dc, fit_tau
    for ig, info in enumerate(comps):
        params[2 + ig * 6: 8 + ig * 6] = [np.float64(v) for v in info[1::2][:6]]
        fit_flags[2 + ig * 6: 8 + ig * 6] = [int(v) for v in info[2::2][:6]]
    if read_only:
        return (modelname, model_code, nu_ref, ngauss, params, fit_flags, alpha, fit_alpha)
    nbin = len(phases)
    if params[1] != 0:
        if P is None:
            print("Need period P for non-zero scattering value TAU.")
            return 0
        params[1] *= nbin / P
    gen = gen_gaussian_portrait_device if device else gen_gaussian_portrait
    model = gen(model_code, params, alpha, phases, freqs, nu_ref)
    if not quiet:
        print("Model Name: %s" % modelname)
        print("Made %d component model with %d profile bins," % (ngauss, nbin))
    return (modelname, ngauss, model)


# ---- TOA output (host text formatting; pplib.py:3380-3503) ----------------------------------
def filter_TOAs(TOAs, flag, cutoff, criterion=">=", pass_unflagged=False, return_culled=False):
    """Keep the TOAs whose attribute ``flag`` satisfies ``<criterion> cutoff`` (pplib.py:3380-3407)."""
    import operator
    ops = {">=": operator.ge, ">": operator.gt, "<=": operator.le, "<": operator.lt,
           "==": operator.eq, "!=": operator.ne}
    test = ops[criterion.strip()]
    kept, culled = [], []
    for toa in TOAs:
        if hasattr(toa, flag):
            (kept if test(getattr(toa, flag), cutoff) else culled).append(toa)
        else:
            (kept if pass_unflagged else culled).append(toa)
    return (kept, culled) if return_culled else kept


def _toa_flag_text(flag, value):
    if hasattr(value, "lower"):
        return " -%s %s" % (flag, value)
    if "int" in str(type(value)):
        return " -%s %d" % (flag, value)
    if "_cov" in flag:
        return " -%s %.1e" % (flag, value)
    if "phs" in flag:
        return " -%s %.8f" % (flag, value)
    if "flux" in flag:
        return " -%s %.5f" % (flag, value)
    return " -%s %.3f" % (flag, value)


def toa_line(toa, inf_is_zero=True):
    """One loosely IPTA-formatted TOA line (pplib.py:3465-3497)."""
    freq = 0.0 if (toa.frequency == np.inf and inf_is_zero) else toa.frequency
    tail = "%.15f   %.3f  %s" % (toa.MJD.fracday(), toa.TOA_error, toa.telescope_code)
    line = "%s %.8f %d" % (toa.archive, freq, toa.MJD.intday()) + tail[1:]
    if toa.DM is not None:
        line += " -pp_dm %.7f" % toa.DM
    if toa.DM_error is not None:
        line += " -pp_dme %.7f" % toa.DM_error
    for flag, value in toa.flags.items():
        if value is not None:
            line += _toa_flag_text(flag, value)
    return line


def write_TOAs(TOAs, inf_is_zero=True, SNR_cutoff=0.0, outfile=None, append=True):
    """Write loosely IPTA-formatted TOAs to ``outfile`` or standard output (pplib.py:3445-3503);
    only TOAs with an ``snr`` flag >= SNR_cutoff are written."""
    toas = TOAs if hasattr(TOAs, "__len__") else [TOAs]
    toas = filter_TOAs(toas, "snr", SNR_cutoff, ">=", pass_unflagged=False)
    lines = [toa_line(t, inf_is_zero) for t in toas]
    if outfile is None:
        for ln in lines:
            print(ln)
        return
    with open(outfile, "a" if append else "w") as fh:
        for ln in lines:
            fh.write(ln + "\n")
