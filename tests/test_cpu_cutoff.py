"""The harmonic cut-off of the device path (include/ppb200.h pp_plan_set_model_cutoff; csrc/kernels.cuh
k_model_cutoff) restated in numpy and checked on the CPU against the untruncated sums the oracle computes: the bound
|delta chi2| / chi2 < 2 eps and the size of the parameter shift, for analytic models at several shapes, and that a
float32-rounded or noisy template keeps every harmonic."""
import numpy as np
import pytest

from oracle import pp_oracle as orc
from tests import synth


def cutoff_groups(model, eps2=1e-20):
    """numpy restatement of k_model_cutoff: per channel the number of leading groups of 16 harmonics kept."""
    nchan, nbin = model.shape
    N = nbin // 2
    NJ = N // 16
    kj_min = min(N, 64) // 16
    pw = np.abs(np.fft.rfft(model, axis=1)) ** 2          # harmonics 0..N
    k = np.arange(N + 1, dtype=np.float64)
    w = k * k * pw
    out = np.zeros(nchan, dtype=int)
    for n in range(nchan):
        grp = [w[n, max(1, 16 * j):16 * j + 16].sum() for j in range(NJ)] + [w[n, N]]   # [NJ] = the Nyquist term
        tot = float(np.sum(grp))
        keep, tail = NJ, grp[NJ]
        if tot > 0 and np.isfinite(tot):
            while keep > kj_min and tail + grp[keep - 1] <= eps2 * tot:
                tail += grp[keep - 1]
                keep -= 1
            if not tail <= eps2 * tot:
                keep = NJ
        out[n] = keep
    return out


def chi2_and_grad(data, model, phi, DM, P, freqs, nu_fit, errs, keep=None):
    """chi2(phi, DM), its gradient and Hessian diagonal from the per-channel sums, harmonics 1..N (or the kept ones)."""
    nchan, nbin = data.shape
    N = nbin // 2
    d = np.fft.rfft(data, axis=1)[:, 1:]
    m = np.fft.rfft(model, axis=1)[:, 1:]
    k = np.arange(1, N + 1, dtype=np.float64)
    sF2 = errs ** 2 * nbin / 2.0
    theta = phi + orc.Dconst * DM / P * (freqs ** -2.0 - nu_fit ** -2.0)
    ph = np.exp(2j * np.pi * np.outer(theta, k))
    X = d * np.conj(m) * ph
    use = np.ones_like(X, dtype=bool)
    if keep is not None:
        for n in range(nchan):
            if keep[n] < N // 16:
                use[n, 16 * keep[n] - 1:] = False         # harmonics k >= 16 keep (k = index + 1), Nyquist included
    C = (X.real * use).sum(1) / sF2
    dC = (-2 * np.pi * k * X.imag * use).sum(1) / sF2
    d2C = (-(2 * np.pi * k) ** 2 * X.real * use).sum(1) / sF2
    S = ((np.abs(m) ** 2) * use).sum(1) / sF2
    Sd = (np.abs(d) ** 2).sum(1) / sF2
    chi2 = float((Sd - C * C / S).sum())
    g = float((-2 * C * dC / S).sum())
    H = float((-2 * (dC * dC + C * d2C) / S).sum())
    return chi2, g, H


@pytest.mark.parametrize("nchan,nbin,nu0,bw", [(32, 2048, 1500., 800.), (16, 4096, 1500., 800.), (32, 1024, 600., 400.)])
def test_cutoff_bound_on_chi2_and_phase(nchan, nbin, nu0, bw):
    freqs, model = synth.example_model(nchan, nbin, nu0, bw)            # float64 as generated
    c = synth.make_case(nchan, nbin, nu0, bw, 4242)
    data, P = c["data"], c["P"]
    errs = orc.get_noise(data, chans=True)
    keep = cutoff_groups(model)
    assert keep.min() >= 4 and keep.mean() < nbin // 32                  # an analytic model: a cut exists (the 400-800 MHz
                                                                         # shape keeps most: its low channels are 2 bins wide)
    nu_fit = freqs.mean()
    full = chi2_and_grad(data, model, c["phi"], c["dDM"], P, freqs, nu_fit, errs)
    cut = chi2_and_grad(data, model, c["phi"], c["dDM"], P, freqs, nu_fit, errs, keep)
    assert abs(cut[0] / full[0] - 1) < 2e-10                              # chi2
    sigma_phi = (2.0 / full[2]) ** 0.5                                    # 1-sigma from H/2
    assert abs(cut[1] - full[1]) / full[2] < 1e-6 * sigma_phi             # Newton shift of phi from the neglected part
    assert abs(cut[2] / full[2] - 1) < 1e-9                               # curvature (error bars)


def test_cutoff_keeps_everything_for_templates_with_a_floor():
    freqs, model = synth.example_model(16, 1024, 1500., 800.)
    NJ = 1024 // 32
    assert (cutoff_groups(model.astype(np.float32).astype(np.float64)) == NJ).all()      # float32 rounding floor
    noisy = model + np.random.RandomState(5).normal(0.0, 1e-4, model.shape)
    assert (cutoff_groups(noisy) == NJ).all()
    assert (cutoff_groups(np.zeros_like(model)) == NJ).all()                              # no power at all
