"""Numerical core of ``ppalign.align_archives`` (ppalign.py:54-243) on the GPU.

Iteratively align and average subints: FFTFIT guess (Ns = nbin) -> phase(+DM)
fit against the current template -> Fourier-domain rotation -> accumulation
weighted by scales/sigma^2 -> new template.  The PSRCHIVE plumbing of the
reference (psradd / vap / psrsmooth shell-outs, writing the output archive) is
out of scope: archives are the ``DataBunch`` objects / ``.npz`` files of
``pptoas.load_data`` and the aligned portrait is returned (and optionally saved
with ``np.save``).
"""
from __future__ import annotations

import numpy as np

from . import pplib
from .pplib import get_plan, _f32, fit_phase_shift, rotate_data, DataBunch  # noqa: F401
from .pptoas import load_data


def normalize_portrait(port, method="rms", weights=None, return_norms=False):
    """Normalise each profile of a portrait (pplib.py:2462-2507)."""
    if method not in ("mean", "max", "prof", "rms", "abs"):
        print("Unknown method for normalize_portrait(...), '%s'." % method)
        return None
    port = np.asarray(port, dtype=np.float64)
    norm_port = np.zeros(port.shape)
    norm_vals = np.ones(len(port))
    good = np.where(port.any(axis=1))[0]
    if method == "prof":
        good_ichans = np.where(port.sum(axis=1) != 0.0)[0]
        w = np.ones(len(good_ichans)) if weights is None else np.asarray(weights)[good_ichans]
        mean_prof = np.average(port[good_ichans], axis=0, weights=w)
        pl = get_plan(1, port.shape[1])
        r = pl.fit_phase_shift_batch(_f32(port[good]), _f32(mean_prof)[None], Ns=100)
        norm_vals[good] = r["scale"]
    elif method == "rms":
        pl = get_plan(1, port.shape[1])
        norm_vals[good] = pl.get_noise_batch(_f32(port[good])[:, None, :])[:, 0]
    elif method == "mean":
        norm_vals[good] = port[good].mean(axis=1)
    elif method == "max":
        norm_vals[good] = port[good].max(axis=1)
    else:
        norm_vals[good] = (port[good] ** 2.0).sum(axis=1) ** 0.5
    norm_port[good] = port[good] / norm_vals[good, None]
    return (norm_port, norm_vals) if return_norms else norm_port


def align_archives(metafile, initial_guess, fit_dm=True, tscrunch=False, pscrunch=True,
                   SNR_cutoff=0.0, outfile=None, norm=None, rot_phase=0.0, place=None, niter=1,
                   quiet=False):
    """Iteratively align and average archives (ppalign.py:54-243).

    metafile: list of archives (DataBunch or .npz path) or a text file of .npz names.
    initial_guess: template portrait [nchan, nbin] (array), or an archive whose first
        subint is used.  Returns DataBunch(port, weights, niter); ``port`` is
        [nchan, nbin] (pscrunch) as in the reference's ``aligned_port[0]``.
    """
    if tscrunch or not pscrunch:
        raise NotImplementedError("tscrunch / Stokes averaging need PSRCHIVE")
    if isinstance(metafile, str):
        datafiles = [ln.strip() for ln in open(metafile, "r").readlines() if ln.strip()]
    else:
        datafiles = list(metafile)
    if isinstance(initial_guess, np.ndarray):
        model_port = np.array(initial_guess, dtype=np.float64)
        model_freqs = None
    else:
        md = initial_guess if isinstance(initial_guess, dict) else load_data(initial_guess)
        model_port = np.array(md.subints[0, 0], dtype=np.float64)
        model_freqs = np.asarray(md.freqs[0], dtype=np.float64)
    nchan, nbin = model_port.shape
    archives = []
    for df in datafiles:
        d = df if isinstance(df, dict) else load_data(df)
        if d.nbin != nbin:
            if not quiet:
                print("%s: %d != %d phase bins.  Skipping it." % (d.filename, d.nbin, nbin))
            continue
        if d.prof_SNR is not None and d.prof_SNR < SNR_cutoff:
            continue
        archives.append(d)
    count = 1
    total_weights = np.zeros(nchan)
    while niter:
        if not quiet:
            print("Doing iteration %d..." % count)
        aligned = np.zeros((nchan, nbin))
        total_weights = np.zeros(nchan)
        for d in archives:
            freqs = np.asarray(d.freqs[0], dtype=np.float64)
            model_ichans = None
            if model_freqs is not None and (len(freqs) != nchan or np.any(freqs != model_freqs)):
                # a different frequency grid than the template: every data channel is fit against
                # (and added to) the template channel closest in frequency (ppalign.py:166-176)
                model_ichans = np.array([np.argmin(abs(model_freqs - f)) for f in freqs])
                if len(np.unique(model_ichans)) != len(model_ichans):
                    # the reference's `aligned_port[ipol, model_ichans] += ...` keeps only the last of
                    # several data channels that share a template channel (numpy fancy-index +=)
                    raise NotImplementedError("several data channels map onto one template channel")
            elif len(freqs) != nchan:
                raise ValueError("%s has %d channels, the template %d" % (d.filename, len(freqs), nchan))
            nchan_d = len(freqs)
            pl = get_plan(nchan_d, nbin)
            pl.set_model(_f32(model_port if model_ichans is None else model_port[model_ichans]), freqs)
            ok_isubs = np.asarray(d.ok_isubs, dtype=int)
            nsub = len(ok_isubs)
            mask = np.zeros((nsub, nchan_d), dtype=np.uint8)
            for i, isub in enumerate(ok_isubs):
                mask[i, np.asarray(d.ok_ichans[isub], dtype=int)] = 1
            subints = _f32(np.asarray(d.subints)[ok_isubs, 0])
            errs = np.ascontiguousarray(np.asarray(d.noise_stds)[ok_isubs, 0], dtype=np.float64)
            snrs = np.ascontiguousarray(np.asarray(d.SNRs)[ok_isubs, 0], dtype=np.float64)
            wts = np.ascontiguousarray(np.asarray(d.weights)[ok_isubs], dtype=np.float64)
            Ps = np.asarray(d.Ps, dtype=np.float64)[ok_isubs]
            DM_guess = float(d.DM) * (not d.dmc)                        # ppalign.py:159
            flags = (1, int(bool(fit_dm)), 0, 0, 0)
            # FFTFIT guess with Ns = nbin (ppalign.py:179-185; the device dedisperses about the
            # mean frequency and transforms the phase to nu_fit = guess_fit_freq, which is the
            # same continuous optimum), then the fit with nu_outs = zero-covariance (190-193)
            # ... and, in the same call, rotated by the fitted phase and DM and added with the
            # weights scales / sigma^2 (ppalign.py:197-208): the archive crosses PCIe and is
            # Fourier transformed once per iteration
            r = pl.fit_batch(subints, Ps, errs=errs, chan_mask=mask, weights=wts, snrs=snrs,
                             DM_guess=np.full(nsub, DM_guess), nu_fit_mode=1,
                             fit_flags=flags, log10_tau=False, Ns=nbin, semantics="full", align=True)
            if model_ichans is None:
                aligned += r["align_sum"]
                total_weights += r["align_wsum"]
            else:
                aligned[model_ichans] += r["align_sum"]
                total_weights[model_ichans] += r["align_wsum"]
        good = total_weights > 0
        aligned[good] /= total_weights[good, None]                      # :210-212
        model_port = aligned
        niter -= 1
        count += 1
    aligned_port = model_port
    if norm in ("mean", "max", "prof", "rms", "abs"):
        aligned_port = normalize_portrait(aligned_port, norm, weights=None)
    if rot_phase:
        aligned_port = rotate_data(aligned_port, rot_phase)
    if place is not None:                                               # :222-226
        prof = np.average(aligned_port, axis=0)
        delta = prof.max() * pplib._wrapped_gaussian(len(prof), place, 0.0001)
        phase = fit_phase_shift(prof, delta, Ns=nbin).phase
        aligned_port = rotate_data(aligned_port, phase)
    if outfile is not None:
        np.save(outfile, aligned_port)
    return DataBunch(port=aligned_port, weights=total_weights, niter=count - 1)
