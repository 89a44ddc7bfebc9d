"""CPU oracle for the PulsePortraiture extended-FFTFIT hot path.

*** TEST INFRASTRUCTURE -- NOT PRODUCT CODE. ***
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``pulseportraiture_b200``) never imports it and has no CPU fallback.

This is a numpy/scipy *restatement* of the reference's algorithm (the
reference, pennucci/PulsePortraiture, is Python 2 and cannot be imported; see
``tests/golden/ref_shim.py``).  It is written vectorised over channels (the
reference loops over channels in Python), so it is a somewhat *faster* CPU
implementation of the same arithmetic than the reference itself.  Every
function cites the reference lines it follows.

Parity pin: the reference ships no tests or golden vectors (SURVEY.md section
4), so this oracle is pinned against outputs of the reference's own functions
executed in the build container through the text shim
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; checked by
``tests/test_oracle_golden.py``).  Third-party arithmetic (numpy pocketfft,
scipy.optimize TNC / trust-ncg / Newton-CG / brute+fmin) is called exactly as
the reference calls it; the versions used to generate the goldens are stored
in each fixture.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.optimize as opt

# pplib.py:44-66
Dconst = 0.000241 ** -1          # "traditional" dispersion constant
F0_fact = 0                      # DC harmonic is ignored in the fits
scattering_alpha = -4.0


class DataBunch(dict):
    """dict with attribute access (pplib.py:125-136)."""

    def __init__(self, **kw):
        dict.__init__(self, kw)
        self.__dict__ = self


# --------------------------------------------------------------------------
# synthetic model portraits (inputs of every parity case)
# --------------------------------------------------------------------------
def get_bin_centers(nbin, lo=0.0, hi=1.0):
    """pplib.py:671-684."""
    half = (hi - lo) / (2.0 * nbin)
    return np.linspace(lo + half, hi - half, nbin)


def read_gmodel(path):
    """Parse a ppgauss ``.gmodel`` file (pplib.py:2867-2925, read-only part).

    Returns dict(name, code, nu_ref, dc, tau, alpha, comps[ngauss,6]).
    comps columns: loc, m_loc, wid, m_wid, amp, m_amp.
    """
    out = dict(name="", code="000", nu_ref=None, dc=0.0, tau=0.0,
               alpha=scattering_alpha, comps=[])
    with open(path, "r") as fh:
        for line in fh:
            tok = line.split()
            if not tok or tok[0].startswith("#"):
                continue
            key = tok[0]
            if key == "MODEL":
                out["name"] = tok[1]
            elif key == "CODE":
                out["code"] = tok[1]
            elif key == "FREQ":
                out["nu_ref"] = float(tok[1])
            elif key == "DC":
                out["dc"] = float(tok[1])
            elif key == "TAU":
                out["tau"] = float(tok[1])
            elif key == "ALPHA":
                out["alpha"] = float(tok[1])
            elif key.startswith("COMP"):
                out["comps"].append([float(v) for v in tok[1::2][:6]])
    out["comps"] = np.array(out["comps"], dtype=np.float64).reshape(-1, 6)
    return out


def _evolve(freqs, nu_ref, ref_vals, evo, code):
    """pplib.py:996-1046: '0' power law, '1' linear."""
    freqs = np.asarray(freqs, dtype=np.float64)
    if code == "0":
        return np.exp(np.outer(np.log(freqs) - np.log(nu_ref), evo) +
                      np.log(ref_vals)[None, :])
    if code == "1":
        return np.outer(freqs - nu_ref, evo) + ref_vals[None, :]
    raise ValueError("unknown evolution code %r" % code)


def _gaussian_rows(nbin, locs, wids):
    """Peak-normalised wrapped Gaussians, one row per (loc, wid) pair.

    Follows pplib.py:770-825 (gaussian_profile, norm=False): bins further than
    20 sigma are zero; the profile is wrapped to within half a turn of the
    mean; the peak is re-normalised so the *continuous* Gaussian centred at
    ``loc`` has unit amplitude.
    """
    locs = np.asarray(locs, dtype=np.float64)
    wids = np.asarray(wids, dtype=np.float64)
    out = np.zeros((locs.size, nbin))
    centers = get_bin_centers(nbin)
    fwhm2sig = 2.0 * np.sqrt(2.0 * np.log(2.0))
    for i, (loc, wid) in enumerate(zip(locs, wids)):
        if not wid > 0.0:
            continue
        sigma = wid / fwhm2sig
        mean = loc % 1.0
        x = centers.copy()
        if mean < 0.5:
            x = np.where(x > mean + 0.5, x - 1.0, x)
        else:
            x = np.where(x < mean - 0.5, x + 1.0, x)
        z = (x - mean) / sigma
        ok = np.fabs(z) < 20.0
        row = np.zeros(nbin)
        row[ok] = np.exp(-0.5 * z[ok] ** 2.0) / (sigma * np.sqrt(2 * np.pi))
        if np.max(np.abs(row)) != 0.0:
            ipk = row.argmax()
            zpk = (x[ipk] - loc) / sigma
            row *= np.exp(-0.5 * zpk ** 2.0) / row[ipk]
        out[i] = row
    return out


def gen_gaussian_portrait(gm, phases, freqs, P=None, tau_override=None):
    """Evolving-Gaussian model portrait (pplib.py:853-930 via read_model
    2926-2936).  ``gm`` is the dict from :func:`read_gmodel`."""
    freqs = np.asarray(freqs, dtype=np.float64)
    nbin = len(phases)
    nchan = len(freqs)
    comps = gm["comps"]
    code = gm["code"]
    locs = _evolve(freqs, gm["nu_ref"], comps[:, 0], comps[:, 1], code[0])
    wids = _evolve(freqs, gm["nu_ref"], comps[:, 2], comps[:, 3], code[1])
    amps = _evolve(freqs, gm["nu_ref"], comps[:, 4], comps[:, 5], code[2])
    port = np.full((nchan, nbin), gm["dc"], dtype=np.float64)
    for ic in range(nchan):
        rows = _gaussian_rows(nbin, locs[ic], wids[ic])
        for ig in range(comps.shape[0]):
            port[ic] += amps[ic, ig] * rows[ig]
    tau = gm["tau"] if tau_override is None else tau_override
    if tau != 0.0:
        if P is None:
            raise ValueError("need P for non-zero TAU")
        tau_bin = tau * nbin / P                      # read_model 2930-2935
        taus = scattering_times(tau_bin / nbin, gm["alpha"], freqs,
                                gm["nu_ref"])
        port = np.fft.irfft(scattering_portrait_FT(taus, nbin) *
                            np.fft.rfft(port, axis=-1), axis=-1)
    return port


def make_freqs(nchan, nu0, bw):
    """Channel centre frequencies as make_fake_pulsar (pplib.py:3236-3240)."""
    return np.linspace(nu0 - bw / 2 + bw / (2.0 * nchan),
                       nu0 + bw / 2 - bw / (2.0 * nchan), nchan)


# --------------------------------------------------------------------------
# L2 signal utilities
# --------------------------------------------------------------------------
def get_noise_PS(data, frac=4, chans=False):
    """Power-spectrum noise estimate (pplib.py:2227-2253)."""
    data = np.asarray(data, dtype=np.float64)
    if chans:
        FT = np.fft.rfft(data, axis=-1)
        pows = (FT.real ** 2 + FT.imag ** 2) / data.shape[-1]
        kc = int((1 - frac ** -1) * pows.shape[-1])
        return np.sqrt(np.mean(pows[:, kc:], axis=-1))
    rav = data.ravel()
    FT = np.fft.rfft(rav)
    pows = (FT.real ** 2 + FT.imag ** 2) / len(rav)
    kc = int((1 - frac ** -1) * len(pows))
    return np.sqrt(np.mean(pows[kc:]))


get_noise = get_noise_PS


def find_kc(pows):
    """Critical harmonic where the noise floor of a power spectrum begins: scipy.optimize.brute fit
    (Ns = 20 per axis, no polish) of b exp(-a k) + dc to log10(pows); the first k with exp(-a k) < 0.005
    (pplib.py:1448-1495, fn = 'exp_dc')."""
    data = np.log10(np.asarray(pows, dtype=np.float64))
    k = np.arange(len(data))

    def chi2(params):
        a, b, dc = params
        return np.sum((data - (b * np.exp(-a * k) + dc)) ** 2.0)
    ranges = [(len(data) ** -1.0, 1.0), (0, data.max() - data.min()), (data.min(), data.max())]
    a, b, dc = opt.brute(chi2, ranges, Ns=20, full_output=False, finish=None)
    hit = np.where(np.exp(-a * k) < 0.005)[0]
    return hit.min() if len(hit) else len(data) - 1


def get_noise_fit(data, fact=1.1, chans=False):
    """Noise from the harmonics above fact * find_kc(power spectrum) (pplib.py:2255-2284)."""
    data = np.asarray(data, dtype=np.float64)

    def one(x):
        FT = np.fft.rfft(x)
        pows = (FT.real ** 2 + FT.imag ** 2) / len(x)
        k_crit = fact * find_kc(pows)
        if k_crit >= len(pows):
            k_crit = min(int(0.99 * len(pows)), k_crit)
        return np.sqrt(np.mean(pows[int(k_crit):]))
    if chans:
        return np.array([one(row) for row in data])
    return one(data.ravel())


def gaussian_profile_FT(nbin, loc, wid, amp):
    """Windowed analytic Fourier transform of a Gaussian profile (pptoaslib.py:14-50)."""
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
    retvals = np.exp(-b ** 2) * (erf(a - b * 1j) + erf(a + b * 1j)) / 2 * (amp * nbin)
    if loc != 0.0:
        retvals = retvals * np.exp(-harmind * 2.0j * np.pi * loc)
    return np.nan_to_num(retvals)


def instrumental_response_FT(nbin, wid=0.0, irf_type='rect'):
    """pptoaslib.py:112-145."""
    nharm = nbin // 2 + 1
    if wid == 0.0:
        return np.ones(nharm)
    if irf_type == 'rect':
        return np.sinc(np.arange(nharm) * wid)
    gp = gaussian_profile_FT(nbin, 0.0, wid, 1.0)
    return gp / gp[0]


def instrumental_response_port_FT(nbin, freqs, DM=0.0, P=1.0, wids=(), irf_types=()):
    """pptoaslib.py:147-179 (the smearing width has no DM factor there: kept)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    nharm = nbin // 2 + 1
    if DM == 0.0 and len(wids) == 0:
        return np.ones([len(freqs), nharm])
    resp = np.ones([len(freqs), nharm], dtype=complex)
    for wid, typ in zip(wids, irf_types):
        resp *= instrumental_response_FT(nbin, wid, typ)[None, :]
    if DM:
        chan_bw = abs(freqs[1] - freqs[0])
        for i, f in enumerate(freqs):
            resp[i] *= instrumental_response_FT(nbin, 8.3e-6 * chan_bw / (f / 1e3) ** 3 / P, 'rect')
    return resp


def add_instrumental_response(model, freqs, DM=0.0, P=1.0, wids=(), irf_types=()):
    """pptoas.py:388-394: modelx = irfft(inst_resp_port_FT * rfft(modelx))."""
    nbin = model.shape[-1]
    resp = instrumental_response_port_FT(nbin, freqs, DM, P, wids, irf_types)
    return np.fft.irfft(resp * np.fft.rfft(model, axis=-1), axis=-1)


def rotate_data(data, phase=0.0, DM=0.0, P=None, freqs=None, nu_ref=np.inf):
    """Fourier-domain rotation / dedispersion of a profile or portrait
    (pplib.py:2338-2426 for 1-D/2-D input; 2428-2460; 2548-2559).
    Positive phase/DM rotate to earlier phase."""
    data = np.asarray(data, dtype=np.float64)
    FT = np.fft.rfft(data, axis=-1)
    k = np.arange(FT.shape[-1])
    if DM == 0.0 or DM is None:
        FT = FT * np.exp(2.0j * np.pi * phase * k)
    else:
        freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
        shifts = phase + (Dconst * DM / P) * (freqs ** -2.0 - nu_ref ** -2.0)
        ph = np.exp(2.0j * np.pi * np.outer(shifts, k))
        FT = FT * (ph if data.ndim == 2 else ph[0])
    return np.fft.irfft(FT, axis=-1)


def phase_transform(phi, DM, nu_ref1=np.inf, nu_ref2=np.inf, P=None,
                    mod=False):
    """pplib.py:2592-2616."""
    if P is None:
        P, mod = 1.0, False
    out = phi + Dconst * DM / P * (nu_ref2 ** -2.0 - nu_ref1 ** -2.0)
    if mod:
        out = np.float64(out)
        if abs(out) >= 0.5:
            out = out % 1
        if out >= 0.5:
            out = out - 1.0
    return out


def guess_fit_freq(freqs, SNRs=None):
    """pplib.py:2618-2632."""
    freqs = np.asarray(freqs, dtype=np.float64)
    nu0 = (freqs.min() + freqs.max()) * 0.5
    if SNRs is None:
        SNRs = np.ones(len(freqs))
    w = SNRs * freqs ** -2
    return nu0 + np.sum((freqs - nu0) * w) / np.sum(w)


def scattering_times(tau, alpha, freqs, nu_tau):
    """pplib.py:4049-4053."""
    return tau * (np.asarray(freqs, dtype=np.float64) / nu_tau) ** alpha


def scattering_portrait_FT(taus, nbin):
    """B_nk = 1/(1 + 2 pi i k tau_n) (pplib.py:4055-4095)."""
    taus = np.atleast_1d(np.asarray(taus, dtype=np.float64))
    k = np.arange(nbin // 2 + 1)
    return 1.0 / (1.0 + 2.0j * np.pi * np.outer(taus, k))


def _spectra(data, model):
    """rfft + DC zeroing shared by all three fits (pplib.py:2127-2130,
    pptoaslib.py:976-979)."""
    dFT = np.fft.rfft(np.asarray(data, dtype=np.float64), axis=-1)
    mFT = np.fft.rfft(np.asarray(model, dtype=np.float64), axis=-1)
    dFT[..., 0] *= F0_fact
    mFT[..., 0] *= F0_fact
    return dFT, mFT


# --------------------------------------------------------------------------
# A3: 1-D FFTFIT  (pplib.py:1244-1280, 2054-2100)
# --------------------------------------------------------------------------
def _pshift_fun(phase, X, err):
    """-Re sum_k X_k e^{2 pi i k phase} / err^2 with X = d conj(m)
    (pplib.py:1244-1256)."""
    k = np.arange(len(X))
    return -np.real((X * np.exp(2.0j * np.pi * k * np.float64(phase))).sum()) \
        / err ** 2.0


def _pshift_fun_ref(phase, mFT, dFT, err):
    # argument order of the reference callback (for opt.brute args=...)
    k = np.arange(len(mFT))
    ph = np.exp(k * 2.0j * np.pi * phase)
    return -np.real((dFT * np.conj(mFT) * ph).sum()) / err ** 2.0


def _pshift_2deriv(phase, X, err):
    """pplib.py:1270-1280."""
    k = np.arange(len(X))
    return -np.real((-4.0 * np.pi ** 2 * k ** 2.0 * X *
                     np.exp(2.0j * np.pi * k * phase)).sum()) / err ** 2.0


def fit_phase_shift_grid(data, model, noise=None, bounds=(-0.5, 0.5), Ns=100):
    """The brute-force stage only: returns (lag_index, grid, values).

    Grid = np.mgrid[lo:hi:Ns*1j] exactly as scipy.optimize.brute builds it
    (pplib.py:2085-2086): Ns points *inclusive* of both ends."""
    dFT, mFT = _spectra(data, model)
    if noise is None:
        err = get_noise_PS(data) * np.sqrt(len(data) / 2.0)
    else:
        err = noise * np.sqrt(len(data) / 2.0)
    grid = np.mgrid[bounds[0]:bounds[1]:complex(0, Ns)]
    vals = np.array([_pshift_fun_ref(g, mFT, dFT, err) for g in grid])
    return int(np.argmin(vals)), grid, vals


def fit_phase_shift(data, model, noise=None, bounds=(-0.5, 0.5), Ns=100,
                    polish="fmin"):
    """1-D FFTFIT (pplib.py:2054-2100).

    polish='fmin' reproduces the reference (scipy brute + Nelder-Mead);
    polish='exact' replaces the polish by a bracketed scalar minimisation to
    1e-14 (the *true* minimiser of the same objective)."""
    data = np.asarray(data, dtype=np.float64)
    dFT, mFT = _spectra(data, model)
    if noise is None:
        err = get_noise_PS(data) * np.sqrt(len(data) / 2.0)
    else:
        err = noise * np.sqrt(len(data) / 2.0)
    d = np.real(np.sum(dFT * np.conj(dFT))) / err ** 2.0
    p = np.real(np.sum(mFT * np.conj(mFT))) / err ** 2.0
    start = time.time()
    if polish == "fmin":
        res = opt.brute(_pshift_fun_ref, [tuple(bounds)],
                        args=(mFT, dFT, err), Ns=Ns, full_output=True)
        phase, fmin = res[0][0], res[1]
        lag = int(np.argmin(res[3]))
    else:
        X = dFT * np.conj(mFT)
        grid = np.mgrid[bounds[0]:bounds[1]:complex(0, Ns)]
        vals = np.array([_pshift_fun(g, X, err) for g in grid])
        lag = int(np.argmin(vals))
        step = grid[1] - grid[0]
        r = opt.minimize_scalar(_pshift_fun, args=(X, err),
                                bracket=(grid[lag] - step, grid[lag],
                                         grid[lag] + step),
                                method="brent", tol=1e-14)
        phase, fmin = r.x, r.fun
    duration = time.time() - start
    X = dFT * np.conj(mFT)
    scale = -fmin / p
    phase_err = (scale * _pshift_2deriv(phase, X, err)) ** -0.5
    scale_err = p ** -0.5
    red_chi2 = (d - (fmin ** 2) / p) / (len(data) - 2)
    snr = pow(scale ** 2 * p, 0.5)
    return DataBunch(phase=phase, phase_err=phase_err, scale=scale,
                     scale_err=scale_err, snr=snr, red_chi2=red_chi2,
                     duration=duration, lag_index=lag)


# --------------------------------------------------------------------------
# A1/A2: phi + DM fit (pplib.py:1282-1391, 2102-2204)
# --------------------------------------------------------------------------
def _chan_sums(phi, DM, X, P, freqs, nu_ref, order=2):
    """Per-channel C, C', C'' of the cross-spectrum X = d conj(m) rotated by
    theta_n = phi + Dconst DM (nu_n^-2 - nu_ref^-2)/P (pplib.py:1315-1322,
    1337-1345, 1372-1381)."""
    k = np.arange(X.shape[1])
    g = (freqs ** -2.0 - nu_ref ** -2.0) * (Dconst / P)
    theta = phi + DM * g
    Z = X * np.exp(2.0j * np.pi * np.outer(theta, k))
    C = Z.real.sum(axis=1)
    out = [C, g]
    if order >= 1:
        out.append((-2.0 * np.pi * k * Z.imag).sum(axis=1))           # C'
    if order >= 2:
        out.append((-(2.0 * np.pi * k) ** 2 * Z.real).sum(axis=1))    # C''
    return out


def fit_portrait_function(params, X, w, P, freqs, nu_ref):
    """f = -sum_n C_n^2/(sigma_n^2 p_n); w = 1/(sigma_F^2 p_n)
    (pplib.py:1282-1325)."""
    C, _ = _chan_sums(params[0], params[1], X, P, freqs, nu_ref, order=0)
    return -(C * C * w).sum()


def fit_portrait_function_deriv(params, X, w, P, freqs, nu_ref):
    """pplib.py:1327-1350."""
    C, g, C1 = _chan_sums(params[0], params[1], X, P, freqs, nu_ref, order=1)
    t = -2.0 * C * C1 * w
    return np.array([t.sum(), (t * g).sum()])


def fit_portrait_function_2deriv(params, X, w, P, freqs, nu_ref):
    """Returns ([H_phiphi, H_DMDM, H_phiDM], nu_zero) (pplib.py:1352-1391)."""
    C, g, C1, C2 = _chan_sums(params[0], params[1], X, P, freqs, nu_ref)
    W = (C1 * C1 + C * C2) * w
    H = -2.0 * np.array([W.sum(), (W * g * g).sum(), (W * g).sum()])
    nu_zero = (W.sum() / np.sum(W * freqs ** -2)) ** 0.5
    return H, nu_zero


def fit_portrait(data, model, init_params, P, freqs, nu_fit=None, nu_out=None,
                 errs=None, bounds=[(None, None), (None, None)], id=None,
                 quiet=True):
    """phi+DM wideband fit with TNC (pplib.py:2102-2204)."""
    data = np.asarray(data, dtype=np.float64)
    freqs = np.asarray(freqs, dtype=np.float64)
    nbin = data.shape[1]
    dFT, mFT = _spectra(data, model)
    if errs is None:
        errsF = get_noise_PS(data, chans=True) * np.sqrt(nbin / 2.0)
    else:
        errsF = np.array(errs, dtype=np.float64) * np.sqrt(nbin / 2.0)
    d = np.sum((dFT.real ** 2 + dFT.imag ** 2) / errsF[:, None] ** 2)
    p_n = np.sum(mFT.real ** 2 + mFT.imag ** 2, axis=1)
    if nu_fit is None:
        nu_fit = freqs.mean()
    X = dFT * np.conj(mFT)
    w = 1.0 / (errsF ** 2 * p_n)
    args = (X, w, P, freqs, nu_fit)
    start = time.time()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")       # TNC ignores 'maxiter' (App. B)
        res = opt.minimize(fit_portrait_function, init_params, args=args,
                           method="TNC", jac=fit_portrait_function_deriv,
                           bounds=bounds,
                           options={"maxiter": 1000, "disp": False,
                                    "xtol": 1e-10})
    duration = time.time() - start
    phi, DM = res.x
    nu_zero = fit_portrait_function_2deriv(res.x, *args)[1]
    if nu_out is None:
        nu_out = nu_zero
    phi_out = phase_transform(phi, DM, nu_fit, nu_out, P, mod=True)
    H3 = fit_portrait_function_2deriv(np.array([phi_out, DM]), X, w, P, freqs,
                                      nu_out)[0]
    H = np.array([[H3[0], H3[2]], [H3[2], H3[1]]])
    cov = np.linalg.inv(0.5 * H)
    param_errs = list(cov.diagonal() ** 0.5)
    dof = data.size - (len(freqs) + 2)
    chi2 = d + res.fun
    C = _chan_sums(phi, DM, X, P, freqs, nu_fit, order=0)[0]
    scales = C / p_n                                    # get_scales 2310-2336
    scale_errs = pow(p_n / errsF ** 2.0, -0.5)
    snr = pow(np.sum(scales ** 2.0 * p_n / errsF ** 2.0), 0.5)
    return DataBunch(phase=phi_out, phase_err=param_errs[0], DM=DM,
                     DM_err=param_errs[1], scales=scales,
                     scale_errs=scale_errs, nu_ref=nu_out,
                     covariance=cov[0, 1], chi2=chi2, red_chi2=chi2 / dof,
                     snr=snr, duration=duration, nfeval=res.nfev,
                     return_code=res.status)


# --------------------------------------------------------------------------
# A4-A8: phi, DM, GM, tau, alpha fit (pptoaslib.py:181-1096)
# --------------------------------------------------------------------------
class _FullProblem(object):
    """Holds the per-subint arrays and evaluates per-channel primitives.

    Everything the reference builds as [2,nchan,nharm] / [2,2,nchan,nharm]
    temporaries (pptoaslib.py:318-523) is restated through the six primitive
    per-channel sums w.r.t. (theta_n, tau_n) plus the chain rule
    (SURVEY Appendix A)."""

    def __init__(self, dFT, mFT, errsF, P, freqs, nu_DM, nu_GM, nu_tau,
                 fit_flags, log10_tau):
        self.X = dFT * np.conj(mFT) / errsF[:, None] ** 2
        self.M = (mFT.real ** 2 + mFT.imag ** 2) / errsF[:, None] ** 2
        self.P = P
        self.freqs = np.asarray(freqs, dtype=np.float64)
        self.nu_DM, self.nu_GM, self.nu_tau = nu_DM, nu_GM, nu_tau
        self.flags = np.array([1.0 if f else 0.0 for f in fit_flags])
        self.log10_tau = log10_tau
        self.k = np.arange(self.X.shape[1])
        self.w = 2.0 * np.pi * self.k

    # ---- per-channel Jacobians -------------------------------------------
    def jac_theta(self):
        """d theta_n / d(phi, DM, GM) (pptoaslib.py:216-225)."""
        f = self.freqs
        return np.array([np.ones(len(f)),
                         Dconst * (f ** -2 - self.nu_DM ** -2) / self.P,
                         Dconst ** 2 * (f ** -4 - self.nu_GM ** -4) / self.P])

    def taus_and_derivs(self, tau, alpha):
        """tau_n and its first/second derivatives w.r.t. the (tau, alpha)
        fit parameters (pptoaslib.py:246-274). ``tau`` is linear here."""
        f = self.freqs
        taus = scattering_times(tau, alpha, f, self.nu_tau)
        lnf = np.log(f / self.nu_tau)
        if not self.log10_tau:
            if taus.sum():
                dt = taus / tau
                dtda = lnf * taus / tau
            else:
                dt = np.zeros(len(f))
                dtda = np.zeros(len(f))
            d2t = np.zeros(len(f))
        else:
            dt = np.log(10.0) * taus
            d2t = np.log(10.0) * dt
            dtda = np.log(10.0) * lnf * taus
        da = lnf * taus
        d2a = lnf * da
        return taus, np.array([dt, da]), np.array([[d2t, dtda], [dtda, d2a]])

    # ---- primitive sums ----------------------------------------------------
    def primitives(self, params, order=2):
        phi, DM, GM, tau, alpha = params
        if self.log10_tau:
            tau = 10 ** tau
        Jth = self.jac_theta()
        theta = phi + DM * Jth[1] + GM * Jth[2]           # 181-214
        taus, Jt, Kt = self.taus_and_derivs(tau, alpha)
        ph = np.exp(2.0j * np.pi * np.outer(theta, self.k))      # 233-238
        scat_on = bool(np.any(taus))
        if scat_on:
            B = 1.0 / (1.0 + 1.0j * np.outer(taus, self.w))
        else:
            B = np.ones(ph.shape, dtype=complex)
        Z = self.X * ph
        r = {}
        r["Jth"], r["Jt"], r["Kt"] = Jth, Jt, Kt
        r["S"] = (np.abs(B) ** 2 * self.M).sum(axis=1)                 # 390
        r["C"] = (Z * np.conj(B)).real.sum(axis=1)                     # 424
        if order >= 1:
            r["Cth"] = (1.0j * self.w * Z * np.conj(B)).real.sum(axis=1)  # 437
            if scat_on and taus.sum():
                dB = -1.0j * self.w * B * B          # = B(B-1)/tau_n (318-330)
                r["Ct"] = (Z * np.conj(dB)).real.sum(axis=1)
                r["St"] = (2.0 * (B * np.conj(dB)).real * self.M).sum(axis=1)
            else:
                dB = None
                r["Ct"] = np.zeros(len(theta))
                r["St"] = np.zeros(len(theta))
        if order >= 2:
            r["Cthth"] = (-(self.w ** 2) * Z * np.conj(B)).real.sum(axis=1)
            if dB is not None:
                d2B = -2.0 * self.w ** 2 * B ** 3   # = 2B(B-1)^2/tau_n^2 (332-356)
                r["Ctt"] = (Z * np.conj(d2B)).real.sum(axis=1)
                r["Ctht"] = (1.0j * self.w * Z * np.conj(dB)).real.sum(axis=1)
                r["Stt"] = (2.0 * (np.abs(dB) ** 2 +
                                   (B * np.conj(d2B)).real) * self.M
                            ).sum(axis=1)
            else:
                z = np.zeros(len(theta))
                r["Ctt"], r["Ctht"], r["Stt"] = z, z.copy(), z.copy()
        return r

    # ---- chain rule to the 5 global parameters ------------------------------
    @staticmethod
    def _first(r):
        Jth, Jt = r["Jth"], r["Jt"]
        dC = np.vstack([r["Cth"] * Jth, r["Ct"] * Jt])            # [5,nchan]
        dS = np.vstack([np.zeros_like(Jth), r["St"] * Jt])
        return dC, dS

    @staticmethod
    def _second(r):
        Jth, Jt, Kt = r["Jth"], r["Jt"], r["Kt"]
        n = Jth.shape[1]
        d2C = np.zeros((5, 5, n))
        d2S = np.zeros((5, 5, n))
        for i in range(3):
            for j in range(3):
                d2C[i, j] = r["Cthth"] * Jth[i] * Jth[j]
            for j in range(2):
                d2C[i, 3 + j] = d2C[3 + j, i] = r["Ctht"] * Jth[i] * Jt[j]
        for i in range(2):
            for j in range(2):
                d2C[3 + i, 3 + j] = r["Ctt"] * Jt[i] * Jt[j] + r["Ct"] * Kt[i, j]
                d2S[3 + i, 3 + j] = r["Stt"] * Jt[i] * Jt[j] + r["St"] * Kt[i, j]
        return d2C, d2S

    # ---- objective, gradient, Hessian ---------------------------------------
    def fun(self, params):
        """pptoaslib.py:525-543."""
        r = self.primitives(params, order=0)
        return -(r["C"] ** 2 / r["S"]).sum()

    def grad(self, params):
        """pptoaslib.py:544-574."""
        r = self.primitives(params, order=1)
        dC, dS = self._first(r)
        C, S = r["C"], r["S"]
        g = -((C ** 2 / S) * (2 * dC / C - dS / S)).sum(axis=-1)
        return g * self.flags

    def hess_per_channel(self, params):
        """pptoaslib.py:576-632 with per_channel=True."""
        r = self.primitives(params, order=2)
        dC, dS = self._first(r)
        d2C, d2S = self._second(r)
        C, S = r["C"], r["S"]
        H = np.zeros((5, 5, len(C)))
        for i in range(5):
            for j in range(5):
                H[i, j] = -2 * (C ** 2 / S) * (
                    d2C[i, j] / C - 0.5 * d2S[i, j] / S +
                    dC[i] * dC[j] / C ** 2 + dS[i] * dS[j] / S ** 2 -
                    (dC[i] * dS[j] + dS[i] * dC[j]) / (C * S)
                ) * self.flags[i] * self.flags[j]
        return H, r

    def hess(self, params):
        return self.hess_per_channel(params)[0].sum(axis=-1)

    def covariance_with_scales(self, params):
        """Lean restatement of pptoaslib.py:645-731 (A, U, S only; the
        reference allocates [5+nchan,5+nchan,nchan]).  Returns the parameter
        covariance (2*UL), the amplitude variances diag(2*LR), the summed
        'A' block, and the scales."""
        r = self.primitives(params, order=2)
        dC, dS = self._first(r)
        d2C, d2S = self._second(r)
        C, S = r["C"], r["S"]
        scales = C / S                                              # 688
        ifit = np.where(self.flags)[0]
        U = (-2.0 * (dC - scales * dS))[ifit]                       # 690, 715
        A = np.zeros((5, 5))
        for i in range(5):
            for j in range(5):
                A[i, j] = (-2 * (C ** 2 / S) *
                           (d2C[i, j] / C - 0.5 * d2S[i, j] / S)
                           ).sum() * self.flags[i] * self.flags[j]  # 694-697
        A = A[np.ix_(ifit, ifit)]
        cinv = 1.0 / (2.0 * S)                                      # 714
        X_inv = np.linalg.inv(A - (U * cinv) @ U.T)                 # 717
        LRdiag = cinv + cinv ** 2 * np.einsum("in,ij,jn->n", U, X_inv, U)
        return 2.0 * X_inv, 2.0 * LRdiag, A, scales                 # 724


def get_nu_zeros(prob, params, option=0):
    """Zero-covariance reference frequencies (pptoaslib.py:733-906).

    Restated without the 0/0-prone divisions by the per-channel Jacobians: the
    reference's ``H21_n = Hij_n[0,1]/phis_deriv[1]`` etc. are the per-channel
    Hessian rows taken with respect to theta_n (resp. ln-frequency) directly.
    """
    flags = [int(bool(f)) for f in prob.flags]
    f = prob.freqs
    nu_DM, nu_GM, nu_tau = prob.nu_DM, prob.nu_GM, prob.nu_tau
    if flags == [1, 1, 1, 1, 1]:                                # 893-901
        sub = _FullProblem.__new__(_FullProblem)
        sub.__dict__.update(prob.__dict__)
        sub.flags = np.array([1.0, 1.0, 0.0, 1.0, 1.0])
        return get_nu_zeros(sub, params, option)
    Hn, r = prob.hess_per_channel(params)
    Jth, Jt = r["Jth"], r["Jt"]
    C, S = r["C"], r["S"]
    a = C ** 2 / S
    dC, dS = prob._first(r)

    def h_theta(j):
        """per-channel Hessian entry (theta_n, param j): Hn[DM, j]/gDM_n."""
        # d2C[theta, j]
        if j < 3:
            d2 = r["Cthth"] * Jth[j]
        else:
            d2 = r["Ctht"] * Jt[j - 3]
        return -2 * a * (d2 / C + r["Cth"] * dC[j] / C ** 2 -
                         (r["Cth"] * dS[j]) / (C * S)) * prob.flags[j]

    def h_lnu(j):
        """per-channel Hessian entry (alpha, param j) / ln(nu_n/nu_tau),
        with the division carried out analytically."""
        tau_lin = 10 ** params[3] if prob.log10_tau else params[3]
        taus = scattering_times(tau_lin, params[4], f, nu_tau)
        # d tau_n/d alpha = lnf * taus ; d2 tau_n/(d alpha d tau) = lnf * k10
        if prob.log10_tau:
            k10 = np.log(10.0) * taus
        elif taus.sum():
            k10 = taus / tau_lin
        else:
            k10 = np.zeros(len(f))
        Ct_a, St_a = r["Ct"] * taus, r["St"] * taus      # d/dalpha / lnf
        if j < 3:
            d2C = r["Ctht"] * Jth[j] * taus
            d2S = 0.0
        elif j == 3:
            d2C = r["Ctt"] * taus * Jt[0] + r["Ct"] * k10
            d2S = r["Stt"] * taus * Jt[0] + r["St"] * k10
        else:
            raise ValueError("alpha-alpha row is never used upstream")
        return -2 * a * (d2C / C - 0.5 * d2S / S + Ct_a * dC[j] / C ** 2 +
                         St_a * dS[j] / S ** 2 -
                         (Ct_a * dS[j] + St_a * dC[j]) / (C * S)
                         ) * prob.flags[j]

    if flags == [1, 1, 0, 0, 0]:                                # 746-752
        h = h_theta(0)
        nu_zero_DM = ((f ** -2 * h).sum() / h.sum()) ** -0.5
        return [nu_zero_DM, nu_GM, nu_tau]
    if flags == [1, 0, 1, 0, 0]:                                # 753-760
        h = h_theta(0)
        nu_zero_GM = ((f ** -4 * h).sum() / h.sum()) ** -0.25
        return [nu_DM, nu_zero_GM, nu_tau]
    if flags == [0, 0, 0, 1, 1]:                                # 761-767
        h = h_lnu(3)
        return [nu_DM, nu_GM, np.exp((np.log(f) * h).sum() / h.sum())]
    if flags == [1, 1, 0, 1, 0]:                                # 768-778
        h21, h23 = h_theta(0), h_theta(3)
        H = Hn.sum(axis=-1)
        H13, H33 = H[3, 0], H[3, 3]
        numer = H13 * (f ** -2 * h23).sum() - H33 * (f ** -2 * h21).sum()
        denom = H13 * h23.sum() - H33 * h21.sum()
        return [(numer / denom) ** -0.5, nu_GM, nu_tau]
    if flags == [1, 1, 0, 1, 1]:                                # 813-836
        h21, h23, h24 = h_theta(0), h_theta(3), h_theta(4)
        h41, h42, h43 = h_lnu(0), h_lnu(1), h_lnu(3)
        H = Hn.sum(axis=-1)
        idx = [0, 1, 3, 4]
        H = H[np.ix_(idx, idx)]
        H11, H22, H33, H44 = np.diag(H)
        H12, H13, H14 = H[0, 1:]
        H23, H24 = H[1, 2:]
        H34 = H[2, 3]
        c1 = (H34 * H34 - H33 * H44)
        c2 = (H13 * H44 - H14 * H34)
        c3 = (H14 * H33 - H13 * H34)
        numer = c1 * (f ** -2 * h21).sum() + c2 * (f ** -2 * h23).sum() + \
            c3 * (f ** -2 * h24).sum()
        denom = c1 * h21.sum() + c2 * h23.sum() + c3 * h24.sum()
        nu_zero_DM = (numer / denom) ** -0.5
        e1 = (H13 * H22 - H12 * H23)
        e2 = (H11 * H23 - H12 * H13)
        e3 = (H12 * H12 - H11 * H22)
        lf = np.log(f)
        numer = e1 * (lf * h41).sum() + e2 * (lf * h42).sum() + \
            e3 * (lf * h43).sum()
        denom = e1 * h41.sum() + e2 * h42.sum() + e3 * h43.sum()
        return [nu_zero_DM, nu_GM, np.exp(numer / denom)]
    if flags == [1, 1, 1, 0, 0] and option in (0, 1):           # 779-812
        # per-channel rows w.r.t. theta_n, split by which Jacobian was
        # divided out in the reference.
        hth = {j: h_theta(j) for j in (0, 1, 2)}
        # Hn[2, j]/gGM_n is the same theta-row (theta second derivs vanish)
        if option == 0:
            H21, H23, H31, H33 = hth[0], hth[2], hth[0], hth[2]
            A_, B_ = (H31 * f ** -4).sum(), H31.sum()
            C_, D_ = (H23 * f ** -2).sum(), H23.sum()
            E_, F_ = (H33 * f ** -4).sum(), H33.sum()
            G_, H_ = (H21 * f ** -2).sum(), H21.sum()
        else:
            H21, H22, H31, H32 = hth[0], hth[1], hth[0], hth[1]
            A_, B_ = (H21 * f ** -4).sum(), H21.sum()
            C_, D_ = (H32 * f ** -2).sum(), H32.sum()
            E_, F_ = (H22 * f ** -4).sum(), H22.sum()
            G_, H_ = (H31 * f ** -2).sum(), H31.sum()
        coeffs = [(A_ * C_ - E_ * G_), 0.0, (E_ * H_ - A_ * D_), 0.0,
                  (F_ * G_ - B_ * C_), 0.0, (B_ * D_ - F_ * H_)]
        roots = np.roots(coeffs)
        roots = np.real(roots[np.where(np.imag(roots) == 0.0)[0]])
        roots = roots[np.where(roots > 0.0)[0]]
        nz = roots[np.argmin(abs(f.mean() - roots))]
        return [nz, nz, nu_tau]
    if flags == [1, 1, 1, 1, 0] and option in (0, 1):           # 837-892
        # Upstream divides by (nu^-2 - nu_DM^-2) and (nu^-4 - nu_GM^-4)
        # *without* the Dconst/P factors (841-842), so the theta-rows pick up
        # those constants; kept as-is ("maybe not right" upstream).
        cD, cG = Dconst / prob.P, Dconst ** 2 / prob.P
        H = Hn.sum(axis=-1)
        H14, H44 = H[3, 0], H[3, 3]
        if option == 0:
            H21, H23, H24 = [cD * h_theta(j) for j in (0, 2, 3)]
            H31, H33, H34 = [cG * h_theta(j) for j in (0, 2, 3)]
            A_, a_ = (f ** -4 * H34).sum(), H34.sum()
            B_, b_ = (f ** -2 * H21).sum(), H21.sum()
            C_, c_ = (f ** -4 * H31).sum(), H31.sum()
            D_, d_ = (f ** -2 * H23).sum(), H23.sum()
            E_, e_ = (f ** -4 * H33).sum(), H33.sum()
            F_, f_ = (f ** -2 * H24).sum(), H24.sum()
            P5 = (A_**2)*B_ + H44*C_*D_ + H14*E_*F_ - H44*B_*E_ - A_*C_*F_ - \
                H14*A_*D_
            P4 = -(A_**2)*b_ - H44*C_*d_ - H14*E_*f_ + H44*b_*E_ + A_*C_*f_ + \
                H14*A_*d_
            P3 = -2*A_*a_*B_ - H44*c_*D_ - H14*e_*F_ + H44*B_*e_ + \
                (A_*c_ + a_*C_)*F_ + H14*a_*D_
            P2 = 2*A_*a_*b_ + H44*c_*d_ + H14*e_*f_ - H44*b_*e_ - \
                (A_*c_ + a_*C_)*f_ - H14*a_*d_
            P1 = (a_**2)*B_ - a_*c_*F_
            P0 = -(a_**2)*b_ + a_*c_*f_
            coeffs = [P5, P4, P3, P2, P1, P0]
        else:
            H21, H22, H24 = [cD * h_theta(j) for j in (0, 1, 3)]
            H31, H32, H34 = [cG * h_theta(j) for j in (0, 1, 3)]
            A_, a_ = (f ** -2 * H24).sum(), H24.sum()
            B_, b_ = (f ** -4 * H31).sum(), H31.sum()
            C_, c_ = (f ** -2 * H21).sum(), H21.sum()
            D_, d_ = (f ** -4 * H32).sum(), H32.sum()
            E_, e_ = (f ** -2 * H22).sum(), H22.sum()
            F_, f_ = (f ** -4 * H34).sum(), H34.sum()
            P4 = (A_**2)*B_ + H44*C_*D_ + H14*E_*F_ - H44*B_*E_ - A_*C_*F_ - \
                H14*A_*D_
            P3 = -2*A_*a_*B_ - H44*c_*D_ - H14*e_*F_ + H44*B_*e_ + \
                (A_*c_ + a_*C_)*F_ + H14*a_*D_
            P2 = -((A_**2)*b_ - (a_**2)*B_) - H44*C_*d_ - H14*E_*f_ + \
                H44*b_*E_ + (A_*C_*f_ - a_*c_*F_) + H14*A_*d_
            P1 = 2*A_*a_*b_ + H44*c_*d_ + H14*e_*f_ - H44*b_*e_ - \
                (A_*c_ + a_*C_)*f_ - H14*a_*d_
            P0 = -(a_**2)*b_ + a_*c_*f_
            coeffs = [P4, P3, P2, P1, P0]
        roots = np.roots(coeffs)
        roots = np.real(roots[np.where(np.imag(roots) == 0.0)[0]])
        roots = roots[np.where(roots > 0.0)[0]]
        roots = roots ** 0.5
        nz = roots[np.argmin(abs(f.mean() - roots))]
        return [nz, nz, nu_tau]
    # every other pattern: reference frequencies unchanged (902-905).
    return [nu_DM, nu_GM, nu_tau]


def fit_portrait_full(data_port, model_port, init_params, P, freqs,
                      nu_fits=[None, None, None], nu_outs=[None, None, None],
                      errs=None, fit_flags=[1, 1, 1, 1, 1],
                      bounds=[(None, None)] * 5, log10_tau=True, option=0,
                      sub_id=None, method="trust-ncg", is_toa=True,
                      quiet=True):
    """5-parameter wideband fit (pptoaslib.py:928-1096)."""
    data_port = np.asarray(data_port, dtype=np.float64)
    freqs = np.asarray(freqs, dtype=np.float64)
    fit_flags = list(fit_flags)
    ifit = np.where(fit_flags)[0]
    nfit = len(ifit)
    dof = data_port.size - (nfit + len(freqs))
    nbin = data_port.shape[-1]
    dFT, mFT = _spectra(data_port, model_port)
    if errs is None:
        errsF = get_noise_PS(data_port, chans=True) * np.sqrt(nbin / 2.0)
    else:
        errsF = np.asarray(errs, dtype=np.float64) * np.sqrt(nbin / 2.0)
    Sd = ((dFT.real ** 2 + dFT.imag ** 2) / errsF[:, None] ** 2).sum()
    nu_fit_DM, nu_fit_GM, nu_fit_tau = [freqs.mean() if v is None else v
                                        for v in nu_fits]
    prob = _FullProblem(dFT, mFT, errsF, P, freqs, nu_fit_DM, nu_fit_GM,
                        nu_fit_tau, fit_flags, log10_tau)
    if method == "trust-ncg":
        kw = dict(hess=prob.hess, options={"gtol": -1})
    elif method == "Newton-CG":
        kw = dict(hess=prob.hess,
                  options={"maxiter": 2000, "disp": False, "xtol": -1})
    elif method == "TNC":
        kw = dict(bounds=bounds,
                  options={"maxiter": 2000, "disp": False, "xtol": 1e-10,
                           "minfev": dof - Sd})
    else:
        raise ValueError("Method '%s' is not implemented." % method)
    start = time.time()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = opt.minimize(prob.fun, np.array(init_params, dtype=np.float64),
                           method=method, jac=prob.grad, **kw)
    duration = time.time() - start
    phi_fit, DM_fit, GM_fit, tau_fit, alpha_fit = res.x
    nu_out_DM, nu_out_GM, nu_out_tau = nu_outs
    if not bool(np.all(nu_outs)):                               # 1041-1047
        nz = get_nu_zeros(prob, res.x, option=option)
        if nu_out_DM is None:
            nu_out_DM = nz[0]
        if nu_out_GM is None:
            nu_out_GM = nz[1]
        if nu_out_tau is None:
            nu_out_tau = nz[2]
    if is_toa:                                                  # 1048-1050
        if fit_flags[1]:
            nu_out_GM = nu_out_DM
        elif fit_flags[2]:
            nu_out_DM = nu_out_GM
    phi_inf = phi_fit - Dconst * DM_fit * nu_fit_DM ** -2 / P - \
        Dconst ** 2 * GM_fit * nu_fit_GM ** -4 / P              # 1052-1053
    phi_out = phi_inf + (Dconst / P) * DM_fit * nu_out_DM ** -2 + \
        (Dconst ** 2 / P) * GM_fit * nu_out_GM ** -4
    if abs(phi_out) >= 0.5:
        phi_out %= 1
    if phi_out >= 0.5:
        phi_out -= 1.0
    tau_lin = 10 ** tau_fit if log10_tau else tau_fit
    tau_out = scattering_times(tau_lin, alpha_fit, nu_out_tau, nu_fit_tau)
    taus = scattering_times(tau_out, alpha_fit, freqs, nu_out_tau)
    if log10_tau:
        tau_out = np.log10(tau_out)
    params = [phi_out, DM_fit, GM_fit, tau_out, alpha_fit]
    prob_out = _FullProblem(dFT, mFT, errsF, P, freqs, nu_out_DM, nu_out_GM,
                            nu_out_tau, fit_flags, log10_tau)
    cov2, var_scales, _, scales = prob_out.covariance_with_scales(params)
    param_errs = np.zeros(5)
    param_errs[ifit] = np.diag(cov2) ** 0.5
    scale_errs = var_scales ** 0.5
    S = (np.abs(scattering_portrait_FT(taus, nbin)) ** 2 *
         (mFT.real ** 2 + mFT.imag ** 2)).sum(axis=-1) / errsF ** 2
    channel_snrs = scales * np.sqrt(S)
    snr = pow(np.sum(channel_snrs ** 2), 0.5)
    chi2 = Sd + res.fun
    return DataBunch(params=params, param_errs=param_errs, phi=phi_out,
                     phi_err=param_errs[0], DM=DM_fit, DM_err=param_errs[1],
                     GM=GM_fit, GM_err=param_errs[2], tau=tau_out,
                     tau_err=param_errs[3], alpha=alpha_fit,
                     alpha_err=param_errs[4], scales=scales,
                     scale_errs=scale_errs, nu_DM=nu_out_DM, nu_GM=nu_out_GM,
                     nu_tau=nu_out_tau, covariance_matrix=cov2, chi2=chi2,
                     red_chi2=chi2 / dof, snr=snr, channel_snrs=channel_snrs,
                     duration=duration, nfeval=res.nfev,
                     return_code=res.status)


# --------------------------------------------------------------------------
# A13: numerical core of GetTOAs.get_TOAs for one subint
# --------------------------------------------------------------------------
def toa_core(port, model, P, freqs, errs, weights=None, SNRs=None,
             DM_stored=0.0, fit_flags=(1, 1, 0, 0, 0), nu_fits=None,
             nu_refs=None, log10_tau=False, tau_guess=0.0, alpha_guess=0.0,
             method="trust-ncg", Ns=100, polish="fmin"):
    """pptoas.py:384-486 re-driven with plain arrays: nu_fit guess,
    dedisperse -> weighted average -> FFTFIT guess -> phase_transform ->
    fit_portrait_full.  Returns (results, phi_guess, nu_fit)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    nbin = port.shape[-1]
    if weights is None:
        weights = np.ones(len(freqs))
    nu_mean = freqs.mean()
    if nu_fits is None:
        nu_fit = guess_fit_freq(freqs, SNRs)                     # 402
        nu_fits = [nu_fit, nu_fit, nu_fit]
    if nu_refs is None:
        nu_refs = [None, None, None]
    rot_port = rotate_data(port, 0.0, DM_stored, P, freqs, nu_mean)   # 422
    rot_prof = np.average(rot_port, axis=0, weights=weights)          # 424
    mprof = np.asarray(model, dtype=np.float64).mean(axis=0)
    if fit_flags[3]:
        B = scattering_portrait_FT(
            np.array([scattering_times(tau_guess, alpha_guess, nu_fits[2],
                                       nu_fits[2])]), nbin)[0]
        mprof = np.fft.irfft(B * np.fft.rfft(mprof))                  # 444-447
    g = fit_phase_shift(rot_prof, mprof, Ns=Ns, polish=polish)        # 448/454
    phi_guess = phase_transform(g.phase, DM_stored, nu_mean, nu_fits[0], P,
                                mod=True)                             # 456
    tg = tau_guess
    if log10_tau:
        if tg == 0.0:
            tg = nbin ** -1
        tg = np.log10(tg)
    guesses = [phi_guess, DM_stored, 0.0, tg, alpha_guess]
    res = fit_portrait_full(port, model, guesses, P, freqs, nu_fits, nu_refs,
                            errs, list(fit_flags), [(None, None)] * 5,
                            log10_tau, option=0, method=method, is_toa=True)
    res.lag_index = g.lag_index
    return res, phi_guess, nu_fits
