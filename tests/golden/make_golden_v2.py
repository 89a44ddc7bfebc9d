"""Generate tests/golden/golden_v2.npz: outputs of the REFERENCE's own functions (ref_shim.py) at the
BASELINE.json shapes and for the round-2 additions.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_v2.py

golden_v1.npz (make_golden.py) stays as it is.  New here:
  b2_*      512 x 2048 phi+DM (config 2 shape): fit_portrait (TNC) and fit_portrait_full (trust-ncg)
  b3_*      256 x 1024 and 512 x 1024 five-parameter fits at sigma = 1.5 (config 3 parity shapes)
  full15_*  every fit_flags pattern of make_golden.py's FULL list again at sigma = 1.5, where
            chi2 and snr^2 are of the same size (the strict 1e-8 chi2 bar)
  ir_*      instrumental_response_FT / _port_FT / gaussian_profile_FT values and a model with the response
  noise_*   get_noise_PS with frac != 4, get_noise_fit
  nb_*      nbin that is not a power of two (1000, 1536, 100): phi+DM, five-parameter, FFTFIT grid, rotation
Inputs are regenerated from the seeds by tests/synth.py; input checksums are stored.
"""
from __future__ import annotations

import io
import os
import sys
import contextlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import numpy as np          # noqa: E402
import scipy                # noqa: E402

from ref_shim import load_reference   # noqa: E402
from tests import synth     # noqa: E402

pl, ptl = load_reference()
out = {}


def put(case, **kw):
    for k, v in kw.items():
        out["%s/%s" % (case, k)] = np.asarray(v)


def bunch_fields(case, prefix, r, fields):
    for f in fields:
        v = r[f]
        if v is None:
            v = np.nan
        out["%s/%s.%s" % (case, prefix, f)] = np.asarray(v, dtype=np.float64)


def quiet_call(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return fn(*a, **k)


FP_FIELDS = ["phase", "phase_err", "DM", "DM_err", "scales", "scale_errs", "nu_ref", "covariance", "chi2",
             "red_chi2", "snr", "nfeval", "return_code"]
FULL_FIELDS = ["params", "param_errs", "phi", "phi_err", "DM", "DM_err", "GM", "GM_err", "tau", "tau_err",
               "alpha", "alpha_err", "scales", "scale_errs", "nu_DM", "nu_GM", "nu_tau", "covariance_matrix",
               "chi2", "red_chi2", "snr", "channel_snrs", "nfeval", "return_code"]
PS_FIELDS = ["phase", "phase_err", "scale", "scale_err", "snr", "red_chi2"]

# ---- config 2 shape: 512 x 2048, phi + DM -------------------------------------------------------------
for seed in (501, 502, 503, 504):
    c = synth.make_case(512, 2048, 1500., 800., seed)
    data, model, freqs, P = c["data"], c["model"], c["freqs"], c["P"]
    case = "b2_%d" % seed
    put(case, cfg=[512, 2048, 1500., 800., seed], in_checksum=synth.checksum(data), truth=[c["phi"], c["dDM"]])
    errs = pl.get_noise(data, chans=True)
    put(case, noise=errs)
    g = pl.fit_phase_shift(data.mean(0), model.mean(0), Ns=100)
    bunch_fields(case, "ps", g, PS_FIELDS)
    r = quiet_call(pl.fit_portrait, data, model, np.array([g.phase, 0.0]), P, freqs, errs=errs)
    bunch_fields(case, "fp", r, FP_FIELDS)
    r = quiet_call(ptl.fit_portrait_full, data, model, [g.phase, 0.0, 0.0, 0.0, 0.0], P, freqs, errs=errs,
                   fit_flags=[1, 1, 0, 0, 0], log10_tau=False)
    bunch_fields(case, "full", r, FULL_FIELDS)
    print(case, "done", flush=True)


def full_case(case, nchan, nbin, nu0, bw, seed, tau_s, flags, log10, option, sigma):
    c = synth.make_case(nchan, nbin, nu0, bw, seed, tau_data_s=tau_s, sigma=sigma)
    data, model, freqs, P = c["data"], c["model"], c["freqs"], c["P"]
    put(case, cfg=[nchan, nbin, nu0, bw, seed, tau_s, log10, option, sigma], flags=flags,
        in_checksum=synth.checksum(data), truth=[c["phi"], c["dDM"]])
    errs = pl.get_noise(data, chans=True)
    tau0 = 0.8 * tau_s / P if tau_s else (0.0 if not flags[3] else 1.0 / nbin)
    tau_init = np.log10(tau0) if (flags[3] and log10) else tau0
    alpha_init = -4.0 if (flags[3] or flags[4]) else 0.0
    g = pl.fit_phase_shift(data.mean(0), model.mean(0), Ns=100)
    phi0 = g.phase if flags[0] else pl.phase_transform(c["phi"], c["dDM"], nu0, freqs.mean(), P, mod=True)
    DM0 = 0.0 if flags[1] else c["dDM"]
    init = [phi0, DM0, 0.0, tau_init, alpha_init]
    put(case, init=init, errs=errs)
    r = quiet_call(ptl.fit_portrait_full, data, model, init, P, freqs, errs=errs, fit_flags=flags,
                   log10_tau=bool(log10), option=option)
    bunch_fields(case, "full", r, FULL_FIELDS)
    print(case, "done", flush=True)


# ---- config 3 parity shapes (nchan in {256, 512} with the verbatim reference) ---------------
for (nchan, seed, flags) in [(256, 601, [1, 1, 0, 1, 1]), (256, 602, [1, 1, 1, 1, 1]),
                             (512, 603, [1, 1, 0, 1, 1]), (512, 604, [1, 1, 1, 1, 1])]:
    full_case("b3_%d" % seed, nchan, 1024, 600., 400., seed, 50e-6, flags, True, 0, 1.5)

# ---- every fit_flags pattern at sigma = 1.5 --------------------------------------------------------------
FULL = [
    (32, 256, 600., 400., 301, 50e-6, [1, 1, 0, 1, 1], True, 0),
    (32, 256, 600., 400., 302, 50e-6, [1, 1, 1, 1, 1], True, 0),
    (32, 256, 600., 400., 303, 50e-6, [1, 1, 0, 1, 0], True, 0),
    (32, 256, 600., 400., 304, 50e-6, [1, 1, 0, 1, 1], False, 0),
    (32, 256, 600., 400., 305, 0.0, [1, 1, 1, 0, 0], False, 0),
    (32, 256, 600., 400., 306, 0.0, [1, 1, 1, 0, 0], False, 1),
    (32, 256, 600., 400., 307, 0.0, [1, 0, 1, 0, 0], False, 0),
    (32, 256, 600., 400., 308, 50e-6, [0, 0, 0, 1, 1], True, 0),
    (32, 256, 600., 400., 309, 50e-6, [1, 1, 1, 1, 0], True, 0),
    (32, 256, 600., 400., 310, 50e-6, [1, 1, 1, 1, 0], True, 1),
    (32, 256, 600., 400., 311, 0.0, [1, 0, 0, 0, 0], False, 0),
    (64, 512, 600., 400., 312, 50e-6, [1, 1, 0, 1, 1], True, 0),
    (64, 512, 600., 400., 313, 50e-6, [1, 1, 1, 1, 1], True, 0),
    (32, 256, 600., 400., 314, 50e-6, [1, 1, 1, 1, 1], False, 0),
]
for (nchan, nbin, nu0, bw, seed, tau_s, flags, log10, option) in FULL:
    full_case("full15_%d" % seed, nchan, nbin, nu0, bw, seed, tau_s, flags, log10, option, 1.5)

# ---- instrumental response (pptoaslib.py:14-50, 112-179; pptoas.py:388-394) ------------------------------
nbin = 256
put("ir", rect=ptl.instrumental_response_FT(nbin, 0.013, 'rect'),
    gauss=np.real(ptl.instrumental_response_FT(nbin, 0.02, 'gauss')),
    gauss_imag_max=np.abs(np.imag(ptl.instrumental_response_FT(nbin, 0.02, 'gauss'))).max(),
    gprof_FT=ptl.gaussian_profile_FT(nbin, 0.3, 0.05, 2.0))
freqs, model = synth.example_model(32, nbin, 1500., 800.)
model = model.astype(np.float32).astype(np.float64)
P = synth.P_EXAMPLE
for tag, (DM, wids, types) in {"wids": (0.0, [0.013, 0.02], ['rect', 'gauss']),
                               "dm": (30.0, [], []),
                               "both": (30.0, [0.01], ['gauss'])}.items():
    resp = ptl.instrumental_response_port_FT(nbin, freqs, DM, P, wids, types)
    conv = np.fft.irfft(resp * np.fft.rfft(model, axis=-1), axis=-1)
    put("ir_" + tag, DM=DM, wids=wids, types=np.array(types, dtype="U8"), resp_real=np.real(resp),
        resp_imag_max=np.abs(np.imag(resp)).max(), model_checksum=synth.checksum(model), conv=conv)
# the response computed on a subset of the channels (freqsx of a subint with zapped channels, pptoas.py:390)
okc = np.array([0, 2, 3, 5, 8, 13, 21, 30])
resp = ptl.instrumental_response_port_FT(nbin, freqs[okc], 30.0, P, [], [])
put("ir_subset", okc=okc, resp_real=np.real(resp))

# ---- noise variants (pplib.py:2206-2284) ------------------------------------------------------------------
c = synth.make_case(16, 512, 1500., 800., 701)
data = c["data"]
put("noise", in_checksum=synth.checksum(data),
    ps_frac8=pl.get_noise_PS(data, frac=8, chans=True), ps_frac2=pl.get_noise_PS(data, frac=2, chans=True),
    ps_frac4=pl.get_noise_PS(data, frac=4, chans=True), ps_frac1=pl.get_noise_PS(data, frac=1, chans=True),
    ps_prof_frac8=pl.get_noise_PS(data[3], frac=8), fit_chans=pl.get_noise_fit(data, chans=True),
    fit_prof=pl.get_noise_fit(data[3]), fit_fact2=pl.get_noise_fit(data, fact=2.0, chans=True))

# ---- nbin that is not a power of two (the reference takes any nbin: np.fft.rfft) ----------------------------
for (nchan, nbin, seed) in [(64, 1000, 801), (32, 1536, 802), (16, 100, 803), (24, 3000, 804)]:
    c = synth.make_case(nchan, nbin, 1500., 800., seed)
    data, model, freqs, P = c["data"], c["model"], c["freqs"], c["P"]
    case = "nb_%d" % seed
    put(case, cfg=[nchan, nbin, 1500., 800., seed], in_checksum=synth.checksum(data), truth=[c["phi"], c["dDM"]])
    errs = pl.get_noise(data, chans=True)
    put(case, noise=errs, rot_row1=pl.rotate_data(data, 0.05, 1e-3, P, freqs, 1400.0)[1])
    g = pl.fit_phase_shift(data.mean(0), model.mean(0), Ns=100)
    bunch_fields(case, "ps", g, PS_FIELDS)
    d1 = np.fft.rfft(data.mean(0)); d1[0] *= 0
    m1 = np.fft.rfft(model.mean(0)); m1[0] *= 0
    err = pl.get_noise(data.mean(0)) * np.sqrt(nbin / 2.0)
    vals = np.array([pl.fit_phase_shift_function(x, m1, d1, err) for x in np.mgrid[-0.5:0.5:100j]])
    put(case, grid_vals=vals, lag=int(np.argmin(vals)))
    r = quiet_call(pl.fit_portrait, data, model, np.array([g.phase, 0.0]), P, freqs, errs=errs)
    bunch_fields(case, "fp", r, FP_FIELDS)
    r = quiet_call(ptl.fit_portrait_full, data, model, [g.phase, 0.0, 0.0, 0.0, 0.0], P, freqs, errs=errs,
                   fit_flags=[1, 1, 0, 0, 0], log10_tau=False)
    bunch_fields(case, "full", r, FULL_FIELDS)
    print(case, "done", flush=True)
full_case("nbfull_811", 32, 1000, 600., 400., 811, 50e-6, [1, 1, 0, 1, 1], True, 0, 1.5)
full_case("nbfull_812", 32, 1536, 600., 400., 812, 50e-6, [1, 1, 1, 1, 1], True, 0, 1.5)

out["meta/versions"] = np.array([np.__version__, scipy.__version__, sys.version.split()[0]])
path = os.path.join(HERE, "golden_v2.npz")
np.savez_compressed(path, **out)
print("wrote %s: %d arrays, %.1f kB" % (path, len(out), os.path.getsize(path) / 1e3))
