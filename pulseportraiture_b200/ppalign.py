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
from .pplib import get_plan, _f32, _mdl, fit_phase_shift, rotate_data, DataBunch  # noqa: F401
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


def _last_wins(model_ichans, okmask):
    """numpy's ``a[idx] += v`` with repeated indices keeps only the LAST contribution per index
    (ppalign.py:205-209 adds the data channels of a subint that way): for every subint, 1 for the last
    usable data channel mapping onto each template channel, 0 for the earlier ones."""
    keep = np.zeros(okmask.shape, dtype=np.float64)
    for s in range(okmask.shape[0]):
        seen = {}
        for c in np.where(okmask[s])[0]:
            seen[int(model_ichans[c])] = c
        keep[s, list(seen.values())] = 1.0
    return keep


def align_archives(metafile, initial_guess, fit_dm=True, tscrunch=False, pscrunch=True,
                   SNR_cutoff=0.0, outfile=None, norm=None, rot_phase=0.0, place=None, niter=1,
                   quiet=False):
    """Iteratively align and average archives (ppalign.py:54-243).

    metafile: list of archives (DataBunch or .npz path) or a text file of .npz names.
    initial_guess: template portrait [nchan, nbin] (array), or an archive whose first
        subint is used.  tscrunch=True pre-averages the subints of every archive
        (pptoas.tscrunch_databunch).  pscrunch=False averages all polarisations of the archives
        ([nsub, npol, nchan, nbin] Stokes subints) with the alignment and weights of the total
        intensity (ppalign.py:203-209).  Returns DataBunch(port, weights, niter); ``port`` is
        [nchan, nbin] (pscrunch, the reference's ``aligned_port[0]``) or [npol, nchan, nbin].
    """
    from .pptoas import tscrunch_databunch
    if isinstance(metafile, str):
        datafiles = [ln.strip() for ln in open(metafile, "r").readlines() if ln.strip()]
    else:
        datafiles = list(metafile)
    model_ok = None
    if isinstance(initial_guess, np.ndarray):
        model_port = np.array(initial_guess, dtype=np.float64)
        model_freqs = None
    else:
        md = initial_guess if isinstance(initial_guess, dict) else load_data(initial_guess)
        model_port = np.array(md.subints[0, 0], dtype=np.float64)
        if md.get("masks") is not None:                                 # ppalign.py:113
            model_port = model_port * np.asarray(md.masks)[0, 0]
        model_freqs = np.asarray(md.freqs[0], dtype=np.float64)
        model_ok = np.asarray(md.ok_ichans[0], dtype=int)
    nchan, nbin = model_port.shape
    npol = 1 if pscrunch else 4
    archives = []
    for df in datafiles:
        d = df if isinstance(df, dict) else load_data(df)
        if d.nbin != nbin:
            if not quiet:
                print("%s: %d != %d phase bins.  Skipping it." % (d.filename, d.nbin, nbin))
            continue
        if not pscrunch and np.asarray(d.subints).shape[1] < npol:
            if not quiet:
                print("%s: has npol = 1.  Skipping it." % d.filename)
            continue
        if d.get("prof_SNR") is not None and d.prof_SNR < SNR_cutoff:
            if not quiet:
                print("%s: %d < %d S/N cutoff.  Skipping it." % (d.filename, d.prof_SNR, SNR_cutoff))
            continue
        if tscrunch:
            d = tscrunch_databunch(d)
        archives.append(d)
    count = 1
    total_weights = np.zeros(nchan)
    aligned = np.zeros((npol, nchan, nbin))
    while niter:
        if not quiet:
            print("Doing iteration %d..." % count)
        aligned = np.zeros((npol, nchan, nbin))
        total_weights = np.zeros(nchan)
        for d in archives:
            freqs = np.asarray(d.freqs[0], dtype=np.float64)
            nchan_d = len(freqs)
            same_freqs = model_freqs is None and nchan_d == nchan or \
                (model_freqs is not None and nchan_d == nchan and not np.any(freqs != model_freqs))
            model_ichans = None
            if not same_freqs:
                if model_freqs is None:
                    raise ValueError("%s has %d channels, the template %d" % (d.filename, nchan_d, nchan))
                # a different frequency grid than the template: every data channel is fit against
                # (and added to) the template channel closest in frequency (ppalign.py:166-176)
                model_ichans = np.array([np.argmin(abs(model_freqs - f)) for f in freqs])
            dup = model_ichans is not None and len(np.unique(model_ichans)) != len(model_ichans)
            pl = get_plan(nchan_d, nbin)
            pl.set_model(_mdl(model_port if model_ichans is None else model_port[model_ichans]), freqs)
            ok_isubs = np.asarray(d.ok_isubs, dtype=int)
            nsub = len(ok_isubs)
            mask = np.zeros((nsub, nchan_d), dtype=np.uint8)
            for i, isub in enumerate(ok_isubs):
                okc = np.asarray(d.ok_ichans[isub], dtype=int)
                if same_freqs and model_ok is not None:                     # ppalign.py:161-163
                    okc = np.intersect1d(okc, model_ok)
                mask[i, okc] = 1
            sub4 = np.asarray(d.subints)
            subints = _f32(sub4[ok_isubs, 0])
            errs = np.ascontiguousarray(np.asarray(d.noise_stds)[ok_isubs, 0], dtype=np.float64)
            snrs = np.ascontiguousarray(np.asarray(d.SNRs)[ok_isubs, 0], dtype=np.float64)
            wts = np.ascontiguousarray(np.asarray(d.weights)[ok_isubs], dtype=np.float64)
            Ps = np.asarray(d.Ps, dtype=np.float64)[ok_isubs]
            DM_guess = float(d.DM) * (not d.dmc)                        # ppalign.py:159
            fused = npol == 1 and not dup
            # FFTFIT guess with Ns = nbin (ppalign.py:179-185; the device dedisperses about the
            # mean frequency and transforms the phase to nu_fit = guess_fit_freq, which is the
            # same continuous optimum), then the fit with nu_outs = zero-covariance (190-193)
            # ... and, in the same call (``fused``), rotated by the fitted phase and DM and added
            # with the weights scales / sigma^2 (ppalign.py:197-208): the archive crosses PCIe and
            # is Fourier transformed once per iteration.  Subints with a single usable channel are
            # fit for the phase only (the 1-channel branch, ppalign.py:194-200).
            nok = mask.sum(axis=1)
            groups = {}
            for i in range(nsub):
                groups.setdefault((1, int(bool(fit_dm)) if nok[i] > 1 else 0, 0, 0, 0), []).append(i)
            align_sum = np.zeros((nchan_d, nbin))
            align_wsum = np.zeros(nchan_d)
            res = {}
            for flags, idx in groups.items():
                idx = np.asarray(idx, dtype=int)
                r = pl.fit_batch(np.ascontiguousarray(subints[idx]), Ps[idx], errs=errs[idx], chan_mask=mask[idx],
                                 weights=wts[idx], snrs=snrs[idx], DM_guess=np.full(len(idx), DM_guess),
                                 nu_fit_mode=1, fit_flags=flags, log10_tau=False, Ns=nbin, semantics="full",
                                 align=fused)
                if fused:
                    align_sum += r["align_sum"]
                    align_wsum += r["align_wsum"]
                for k in ("params", "nu_out", "scales", "return_code"):
                    res.setdefault(k, [None] * nsub)
                    for j, i in enumerate(idx):
                        res[k][i] = r[k][j]
            if fused:
                sums = [align_sum]
            else:
                # Stokes data and / or several data channels per template channel: the weights of the
                # total-intensity fit applied to every polarisation in a second pass over the data
                params = np.array(res["params"])
                nu_ref = np.array(res["nu_out"])[:, 0]
                w = np.array(res["scales"]) / np.where(errs > 0, errs, np.inf) ** 2 * mask
                w[np.array(res["return_code"]) == 3] = 0.0
                if dup:
                    w = w * _last_wins(model_ichans, mask.astype(bool))
                sums = []
                for ipol in range(npol):
                    asum, align_wsum = pl.align_accumulate(_f32(sub4[ok_isubs, ipol]), params[:, 0], params[:, 1],
                                                           Ps, nu_ref, np.ascontiguousarray(w))
                    sums.append(asum)
            for ipol in range(npol):
                if model_ichans is None:
                    aligned[ipol] += sums[ipol]
                else:
                    np.add.at(aligned[ipol], model_ichans, sums[ipol])
            if model_ichans is None:
                total_weights += align_wsum
            else:
                np.add.at(total_weights, model_ichans, align_wsum)
        good = total_weights > 0
        aligned[:, good] /= total_weights[good][None, :, None]              # :210-212
        model_port = aligned[0]
        niter -= 1
        count += 1
    aligned_port = aligned if npol > 1 else model_port
    if norm in ("mean", "max", "prof", "rms", "abs"):
        if npol > 1:
            aligned_port = np.array([normalize_portrait(a, norm, weights=None) for a in aligned_port])
        else:
            aligned_port = normalize_portrait(aligned_port, norm, weights=None)
    rot = (lambda a, ph: np.array([rotate_data(x, ph) for x in a])) if npol > 1 else rotate_data
    if rot_phase:
        aligned_port = rot(aligned_port, rot_phase)
    if place is not None:                                               # :222-226
        prof = np.average(aligned_port[0] if npol > 1 else aligned_port, axis=0)
        delta = prof.max() * pplib._wrapped_gaussian(len(prof), place, 0.0001)
        phase = fit_phase_shift(prof, delta, Ns=nbin).phase
        aligned_port = rot(aligned_port, phase)
    if outfile is not None:
        np.save(outfile, aligned_port)
    return DataBunch(port=aligned_port, weights=total_weights, niter=count - 1)
