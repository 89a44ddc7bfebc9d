"""Seeded synthetic portraits for the parity tests.

Inputs are built with the ORACLE's model generator and rotation (tests may use
the oracle) from the ``example.gmodel`` fixture, exactly as the golden
generator did, so that ``tests/golden/*.npz`` (outputs of the reference's own
functions on these inputs) can be compared without storing the inputs.
"""
from __future__ import annotations

import os

import numpy as np

from oracle import pp_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GMODEL = os.path.join(HERE, "golden", "example.gmodel")
P_EXAMPLE = 1.0 / 345.67890123456789          # examples/example.par:4

_model_cache = {}


def example_model(nchan, nbin, nu0, bw, tau_s=0.0, alpha=-4.0, P=P_EXAMPLE):
    """(freqs, model[nchan,nbin] float64).  tau_s: scattering time [s] at the
    .gmodel reference frequency (0 = unscattered model)."""
    key = (nchan, nbin, nu0, bw, tau_s, alpha, P)
    if key not in _model_cache:
        gm = orc.read_gmodel(GMODEL)
        gm = dict(gm)
        gm["alpha"] = alpha
        freqs = orc.make_freqs(nchan, nu0, bw)
        phases = orc.get_bin_centers(nbin)
        model = orc.gen_gaussian_portrait(gm, phases, freqs, P=P,
                                          tau_override=tau_s)
        _model_cache[key] = (freqs, model)
    f, m = _model_cache[key]
    return f.copy(), m.copy()


def make_case(nchan, nbin, nu0, bw, seed, phi=None, dDM=None, sigma=1.5,
              tau_data_s=0.0, alpha=-4.0, P=P_EXAMPLE, legacy_seed=False,
              scales=None):
    """One synthetic subint.

    data = rotate(model_scattered, -phi, -dDM) + N(0, sigma^2), rounded to
    float32 (the device input type) and returned upcast to float64 so the
    oracle / reference see *the same numbers* as the GPU.
    Returns dict(data, model, freqs, P, phi, dDM, nu0).
    """
    freqs, model = example_model(nchan, nbin, nu0, bw, 0.0, alpha, P)
    if tau_data_s:
        # scatter the unscattered model: tau_n = (tau/P)(nu_n/nu0)^alpha [rot]
        taus = orc.scattering_times(tau_data_s / P, alpha, freqs, nu0)
        src = np.fft.irfft(orc.scattering_portrait_FT(taus, nbin) *
                           np.fft.rfft(model, axis=-1), axis=-1)
    else:
        src = model
    if legacy_seed:
        np.random.seed(seed)
        rng = np.random
    else:
        rng = np.random.RandomState(seed)
    if phi is None:
        phi = rng.uniform(-0.5, 0.5)
    if dDM is None:
        dDM = rng.normal(3e-4, 2e-4)
    if scales is not None:
        src = src * np.asarray(scales)[:, None]
    clean = orc.rotate_data(src, -phi, -dDM, P, freqs, nu0)
    noise = rng.normal(0.0, sigma, clean.shape)
    data = (clean + noise).astype(np.float32).astype(np.float64)
    # the device takes the model as float32 too; everyone sees those values
    model = model.astype(np.float32).astype(np.float64)
    return dict(data=data, model=model, freqs=freqs, P=P, phi=phi, dDM=dDM,
                nu0=nu0)


def checksum(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), (a * a).sum(),
                     (a.ravel() * np.arange(1, a.size + 1)).sum()])
